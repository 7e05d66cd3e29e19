// multiverso-b200 :: K8, LogisticRegression kernels.
//
// Reference hot loops: Dot (Applications/LogisticRegression/src/util/common.h:23-41),
// Objective::Predict / Gradient for linear, sigmoid, softmax (objective/objective.cpp:
// 37-47, 64-100, 113-120, 142-233), the minibatch average in Model::Update
// (model/model.cpp:78-104), regularisers (regular/regular.cpp:21-56) and the FTRL
// objective/updater (objective.cpp:238-341, updater/updater.cpp:80-101).
//
// Layout: W is [out x dim] row-major, dim includes the bias column (the reader
// appends a constant-1 feature as the reference does, input_size += 1).
//   phase 1 (warp per sample): logits -> sigma / softmax -> loss, accuracy,
//            err[i][c] = (p_c - y_c) * weight_i / n
//   phase 2: sparse: warp per sample scatters err * x into grad with red.add
//            dense : grad[c][j] = sum_i err[i][c] x[i][j]  (column-parallel,
//            err broadcast from shared memory, no atomics inside a sample slab)
#include "mvb_common.cuh"

namespace {

constexpr int kMaxOut = 64;

MVB_DEVINL float sigm(float x) { return 1.f / (1.f + __expf(-x)); }

// turn logits (in smem, `out` values) into err/pred/loss for one sample; lane 0 only
MVB_DEVINL void finish_sample(float* lg, int out, int objective, float label, float wgt,
                              float inv_n, float* err, float* pred, float& loss, int& correct) {
  if (objective == 2 && out > 1) {  // softmax, max-subtracted (objective.cpp:202-218)
    float mx = lg[0];
    int arg = 0;
    for (int c = 1; c < out; ++c) if (lg[c] > mx) { mx = lg[c]; arg = c; }
    float sum = 0.f;
    for (int c = 0; c < out; ++c) { lg[c] = __expf(lg[c] - mx); sum += lg[c]; }
    float inv = 1.f / sum;
    int y = (int)label;
    for (int c = 0; c < out; ++c) {
      float p = lg[c] * inv;
      if (pred) pred[c] = p;
      err[c] = (p - (c == y ? 1.f : 0.f)) * wgt * inv_n;
      if (c == y) loss += -__logf(fmaxf(p, 1e-30f)) * wgt;
    }
    correct += (arg == y);
  } else {
    for (int c = 0; c < out; ++c) {
      float yc = (out == 1) ? label : ((int)label == c ? 1.f : 0.f);
      float p = lg[c];
      if (objective >= 1) {  // sigmoid
        p = sigm(p);
        loss += -(yc * __logf(fmaxf(p, 1e-30f)) + (1.f - yc) * __logf(fmaxf(1.f - p, 1e-30f))) * wgt;
      } else {               // linear, squared loss
        loss += 0.5f * (p - yc) * (p - yc) * wgt;
      }
      if (pred) pred[c] = p;
      err[c] = (p - yc) * wgt * inv_n;
      if (out == 1) correct += (objective >= 1 ? ((p > 0.5f) == (yc > 0.5f)) : (fabsf(p - yc) < 0.5f));
    }
    if (out > 1) {
      int arg = 0;
      for (int c = 1; c < out; ++c) if (lg[c] > lg[arg]) arg = c;
      correct += (arg == (int)label);
    }
  }
}

__global__ void __launch_bounds__(128)
lr_sparse_fwd_kernel(MvbLrSparse a) {
  __shared__ float lgs[4][kMaxOut];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float loss = 0.f;
  int correct = 0;
  const float inv_n = 1.f / (float)a.n;
  for (int64_t i = warp; i < a.n; i += nwarps) {
    const int64_t b = a.row_ptr[i], e = a.row_ptr[i + 1];
    for (int c = 0; c < a.out; ++c) {
      const float* wc = a.w + (int64_t)c * a.dim;
      float s = 0.f;
      for (int64_t j = b + lane; j < e; j += 32) {
        int64_t k = a.keys[j];
        float v = a.vals ? a.vals[j] : 1.f;
        if (k >= 0 && k < a.dim) s = fmaf(__ldg(wc + k), v, s);
      }
      s = warp_sum(s);
      if (lane == 0) lgs[wib][c] = s;
    }
    __syncwarp();
    if (lane == 0)
      finish_sample(lgs[wib], a.out, a.objective, a.labels[i], a.sample_w ? a.sample_w[i] : 1.f,
                    inv_n, a.err + i * a.out, a.pred ? a.pred + i * a.out : nullptr, loss, correct);
    __syncwarp();
    if (a.compute_grad) {
      for (int c = 0; c < a.out; ++c) {
        const float ec = a.err[i * a.out + c];
        float* gc = a.grad + (int64_t)c * a.dim;
        for (int64_t j = b + lane; j < e; j += 32) {
          int64_t k = a.keys[j];
          float v = a.vals ? a.vals[j] : 1.f;
          if (k >= 0 && k < a.dim) red_add_f32(gc + k, ec * v);
        }
      }
    }
  }
  if (lane == 0) {
    if (a.loss_sum && loss != 0.f) atomicAdd(a.loss_sum, loss);
    if (a.correct && correct) atomicAdd(a.correct, correct);
  }
}

__global__ void __launch_bounds__(128)
lr_dense_fwd_kernel(MvbLrDense a) {
  __shared__ float lgs[4][kMaxOut];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float loss = 0.f;
  int correct = 0;
  const float inv_n = 1.f / (float)a.n;
  for (int64_t i = warp; i < a.n; i += nwarps) {
    const float* x = a.x + i * a.dim;
    for (int c = 0; c < a.out; ++c) {
      const float* wc = a.w + (int64_t)c * a.dim;
      float s = 0.f;
      for (int64_t j = lane; j < a.dim; j += 32) s = fmaf(__ldg(wc + j), x[j], s);
      s = warp_sum(s);
      if (lane == 0) lgs[wib][c] = s;
    }
    __syncwarp();
    if (lane == 0)
      finish_sample(lgs[wib], a.out, a.objective, a.labels[i], 1.f, inv_n, a.err + i * a.out,
                    a.pred ? a.pred + i * a.out : nullptr, loss, correct);
    __syncwarp();
  }
  if (lane == 0) {
    if (a.loss_sum && loss != 0.f) atomicAdd(a.loss_sum, loss);
    if (a.correct && correct) atomicAdd(a.correct, correct);
  }
}

// grad[c][j] += sum over a slab of samples of err[i][c] * x[i][j]
constexpr int kSlab = 128;
__global__ void __launch_bounds__(128)
lr_dense_grad_kernel(MvbLrDense a) {
  __shared__ float es[kSlab][kMaxOut / 4 + 1];  // up to 16 classes per pass
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i0 = (int64_t)blockIdx.y * kSlab;
  const int cnt = (int)((a.n - i0) < kSlab ? (a.n - i0) : kSlab);
  for (int c0 = 0; c0 < a.out; c0 += 16) {
    const int nc = a.out - c0 < 16 ? a.out - c0 : 16;
    __syncthreads();
    for (int t = threadIdx.x; t < cnt * nc; t += blockDim.x)
      es[t / nc][t % nc] = a.err[(i0 + t / nc) * a.out + c0 + t % nc];
    __syncthreads();
    if (j < a.dim) {
      float acc[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = 0.f;
      for (int i = 0; i < cnt; ++i) {
        float xv = a.x[(i0 + i) * a.dim + j];
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (c < nc) acc[c] = fmaf(es[i][c], xv, acc[c]);
      }
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (c < nc) red_add_f32(a.grad + (int64_t)(c0 + c) * a.dim + j, acc[c]);
    }
  }
}

__global__ void ftrl_weights_kernel(const float* z, const float* n, float* w, int64_t len,
                                    float alpha, float beta, float l1, float l2) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    float zi = z[i];
    float sgn = zi < 0.f ? -1.f : 1.f;
    float out = 0.f;
    if (sgn * zi > l1) out = (sgn * l1 - zi) / ((beta + sqrtf(n[i])) / alpha + l2);
    w[i] = out;
  }
}
__global__ void ftrl_update_kernel(float* z, float* n, const float* w, const float* g, int64_t len,
                                   float alpha) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    float gi = g[i];
    if (gi == 0.f) continue;
    float ni = n[i];
    float sigma = (sqrtf(ni + gi * gi) - sqrtf(ni)) / alpha;
    z[i] += gi - sigma * w[i];
    n[i] = ni + gi * gi;
  }
}
__global__ void regularize_kernel(float* grad, const float* w, int64_t len, int type, float coef) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    float wi = w[i];
    if (type == 1) grad[i] += (wi > 0.f ? coef : (wi < 0.f ? -coef : 0.f));
    else if (type == 2) grad[i] += wi * coef;  // true L2 (reference uses |w|*c)
  }
}

int ew_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  int64_t cap = (int64_t)mvb_num_sms() * 8;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" int mvb_lr_sparse_fwd_bwd(const MvbLrSparse* a, void* stream) {
  if (a->n <= 0) return 0;
  if (a->out > kMaxOut || a->out < 1) return -8;
  int64_t blocks = (a->n + 3) / 4;
  int64_t cap = (int64_t)mvb_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  lr_sparse_fwd_kernel<<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(*a);
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int mvb_lr_dense_fwd_bwd(const MvbLrDense* a, void* stream) {
  if (a->n <= 0) return 0;
  if (a->out > kMaxOut || a->out < 1) return -8;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = (a->n + 3) / 4;
  int64_t cap = (int64_t)mvb_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  lr_dense_fwd_kernel<<<(int)blocks, 128, 0, st>>>(*a);
  if (a->compute_grad) {
    dim3 grid((unsigned)((a->dim + 127) / 128), (unsigned)((a->n + kSlab - 1) / kSlab));
    lr_dense_grad_kernel<<<grid, 128, 0, st>>>(*a);
  }
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int mvb_ftrl_weights(const float* z, const float* n, float* w, int64_t len, float alpha,
                                float beta, float l1, float l2, void* stream) {
  ftrl_weights_kernel<<<ew_blocks(len), 256, 0, (cudaStream_t)stream>>>(z, n, w, len, alpha, beta, l1, l2);
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int mvb_ftrl_update(float* z, float* n, const float* w, const float* g, int64_t len,
                               float alpha, void* stream) {
  ftrl_update_kernel<<<ew_blocks(len), 256, 0, (cudaStream_t)stream>>>(z, n, w, g, len, alpha);
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int mvb_regularize(float* grad, const float* w, int64_t len, int type, float coef,
                              void* stream) {
  regularize_kernel<<<ew_blocks(len), 256, 0, (cudaStream_t)stream>>>(grad, w, len, type, coef);
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
