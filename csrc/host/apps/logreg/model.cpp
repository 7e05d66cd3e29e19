#include "model.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "multiverso/io/io.h"
#include "multiverso/multiverso.h"
#include "multiverso/util/log.h"
#include "multiverso/util/timer.h"

namespace logreg {

using multiverso::Log;
using multiverso::Timer;

namespace {
constexpr float kTiny = 1e-30f;
inline float Sigmoid(float x) { return 1.0f / (1.0f + std::exp(-x)); }
inline float Sign(float x) { return x > 0 ? 1.0f : (x < 0 ? -1.0f : 0.0f); }

ObjectiveKind ParseObjective(const Configure& c) {
  if (c.objective_type == "softmax" && c.output_size > 1) return ObjectiveKind::Softmax;
  if (c.objective_type == "sigmoid" || c.objective_type == "softmax" || c.ftrl()) return ObjectiveKind::Sigmoid;
  return ObjectiveKind::Linear;
}
RegularKind ParseRegular(const std::string& t) {
  if (t == "L1" || t == "l1") return RegularKind::L1;
  if (t == "L2" || t == "l2") return RegularKind::L2;
  return RegularKind::None;
}
}  // namespace

void MiniBatch::Reserve(int64_t max_samples, int64_t max_nnz) {
  row_ptr.resize(max_samples + 1);
  labels.resize(max_samples);
  weights.resize(max_samples);
  keys.resize(max_nnz);
  vals.resize(max_nnz);
}

// ------------------------------------------------------------------------------------ Model
Model::Model(const Configure& config)
    : cfg_(config), dim_(config.input_size + 1), out_(std::max(1, config.output_size)),
      size_(static_cast<int64_t>(out_) * dim_), objective_(ParseObjective(config)),
      regular_(ParseRegular(config.regular_type)), ftrl_(config.ftrl()),
      lr_(config.updater_type == "sgd" ? static_cast<float>(config.learning_rate) : 1.0f) {
  w_.assign(size_, 0.0f);
  delta_.assign(size_, 0.0f);
  touched_mark_.assign(size_, 0);
  if (ftrl_) {
    z_.assign(size_, 0.0f);
    n_.assign(size_, 0.0f);
    delta_n_.assign(size_, 0.0f);
  }
}

std::unique_ptr<Model> Model::Create(const Configure& config) {
  if (config.use_ps) return std::make_unique<PSModel>(config);
  return std::make_unique<Model>(config);
}

void Model::ForwardSample(const MiniBatch& b, int64_t i, float* logits) const {
  const int64_t lo = b.row_ptr[i], hi = b.row_ptr[i + 1];
  for (int c = 0; c < out_; ++c) {
    const float* wc = w_.data() + static_cast<int64_t>(c) * dim_;
    float s = 0;
    for (int64_t j = lo; j < hi; ++j) s += wc[b.keys[j]] * b.vals[j];
    logits[c] = s;
  }
}

// Turns logits into predictions in place, fills err[c] = p_c - y_c and adds the sample's loss /
// correctness to the accumulators.
void Model::FinishSample(float* lg, float label, float* err, double* loss, int64_t* correct) const {
  if (objective_ == ObjectiveKind::Softmax) {
    int arg = 0;
    for (int c = 1; c < out_; ++c)
      if (lg[c] > lg[arg]) arg = c;
    const float mx = lg[arg];
    float sum = 0;
    for (int c = 0; c < out_; ++c) sum += (lg[c] = std::exp(lg[c] - mx));
    const int y = static_cast<int>(label);
    for (int c = 0; c < out_; ++c) {
      lg[c] /= sum;
      err[c] = lg[c] - (c == y ? 1.0f : 0.0f);
    }
    if (y >= 0 && y < out_) *loss -= std::log(std::max(lg[y], kTiny));
    *correct += (arg == y);
    return;
  }
  int arg = 0;
  for (int c = 0; c < out_; ++c) {
    const float yc = out_ == 1 ? label : (static_cast<int>(label) == c ? 1.0f : 0.0f);
    float p = lg[c];
    if (objective_ == ObjectiveKind::Sigmoid) {
      p = Sigmoid(p);
      *loss -= yc * std::log(std::max(p, kTiny)) + (1.0f - yc) * std::log(std::max(1.0f - p, kTiny));
    } else {
      *loss += 0.5 * (p - yc) * (p - yc);
    }
    lg[c] = p;
    err[c] = p - yc;
    if (p > lg[arg]) arg = c;
  }
  if (out_ == 1)
    *correct += objective_ == ObjectiveKind::Sigmoid ? ((lg[0] > 0.5f) == (label > 0.5f)) : (std::fabs(lg[0] - label) < 0.5f);
  else
    *correct += (arg == static_cast<int>(label));
}

// FTRL-proximal closed form (objective.cpp:260-336): w = 0 inside the L1 ball, else
// -(z - sign(z) lambda1) / ((beta + sqrt(n)) / alpha + lambda2)
void Model::RefreshFtrlWeight(int64_t key) {
  const float z = z_[key];
  if (std::fabs(z) <= cfg_.lambda1) {
    w_[key] = 0;
  } else {
    w_[key] = static_cast<float>(-(z - Sign(z) * cfg_.lambda1) /
                                 ((cfg_.beta + std::sqrt(n_[key])) / cfg_.alpha + cfg_.lambda2));
  }
}

void Model::ClearDelta() {
  for (int64_t k : touched_) {
    delta_[k] = 0;
    touched_mark_[k] = 0;
    if (ftrl_) delta_n_[k] = 0;
  }
  touched_.clear();
}

BatchResult Model::Update(const MiniBatch& b) {
  Timer timer;
  BatchResult r;
  if (b.n == 0) return r;
  std::vector<float> logits(out_), err(out_);
  const int64_t nnz = b.row_ptr[b.n];
  // every (class, feature) pair of the batch receives a gradient
  for (int64_t j = 0; j < nnz; ++j)
    for (int c = 0; c < out_; ++c) {
      const int64_t key = static_cast<int64_t>(c) * dim_ + b.keys[j];
      if (!touched_mark_[key]) {
        Touch(key);
        if (ftrl_) RefreshFtrlWeight(key);
      }
    }
  for (int64_t i = 0; i < b.n; ++i) {
    ForwardSample(b, i, logits.data());
    FinishSample(logits.data(), b.labels[i], err.data(), &r.loss, &r.correct);
    const float sw = cfg_.reader_type == "weight" ? b.weights[i] : 1.0f;
    for (int c = 0; c < out_; ++c) {
      float* gc = delta_.data() + static_cast<int64_t>(c) * dim_;
      const float e = err[c] * sw;
      for (int64_t j = b.row_ptr[i]; j < b.row_ptr[i + 1]; ++j) gc[b.keys[j]] += e * b.vals[j];
    }
  }
  const float inv_n = 1.0f / static_cast<float>(b.n);
  const float coef = static_cast<float>(cfg_.regular_coef);
  for (int64_t k : touched_) {
    float g = delta_[k] * inv_n;                  // minibatch average (model.cpp:78-104)
    if (ftrl_) {
      const float sigma = (std::sqrt(n_[k] + g * g) - std::sqrt(n_[k])) / static_cast<float>(cfg_.alpha);
      delta_[k] = -(g - sigma * w_[k]);           // the updater subtracts: z += g - sigma w
      delta_n_[k] = -(g * g);                     //                        n += g^2
      continue;
    }
    if (regular_ == RegularKind::L1) g += coef * Sign(w_[k]);
    else if (regular_ == RegularKind::L2) g += coef * w_[k];   // true L2 (the reference uses |w|)
    delta_[k] = g * lr_;
  }
  compute_ms_ += timer.elapse();
  ApplyDelta();
  ClearDelta();
  ++updates_;
  if (cfg_.updater_type == "sgd")               // SGDUpdater::Process (updater.cpp:44-71)
    lr_ = static_cast<float>(std::max(1e-3, cfg_.learning_rate - updates_ / (cfg_.learning_rate_coef * cfg_.minibatch_size)));
  return r;
}

void Model::ApplyDelta() {
  for (int64_t k : touched_) {
    if (ftrl_) {
      z_[k] -= delta_[k];
      n_[k] -= delta_n_[k];
    } else {
      w_[k] -= delta_[k];
    }
  }
}

BatchResult Model::Predict(const MiniBatch& b, std::vector<float>* predictions) {
  BatchResult r;
  predictions->resize(static_cast<size_t>(b.n) * out_);
  std::vector<float> err(out_);
  if (ftrl_) {
    const int64_t nnz = b.n ? b.row_ptr[b.n] : 0;
    for (int64_t j = 0; j < nnz; ++j)
      for (int c = 0; c < out_; ++c) RefreshFtrlWeight(static_cast<int64_t>(c) * dim_ + b.keys[j]);
  }
  for (int64_t i = 0; i < b.n; ++i) {
    float* p = predictions->data() + static_cast<size_t>(i) * out_;
    ForwardSample(b, i, p);
    FinishSample(p, b.labels[i], err.data(), &r.loss, &r.correct);
  }
  return r;
}

// Model files: dense = raw dump of the weights; sparse = count, then (size_t key, value) per
// non-zero entry; FTRL values are {z, n} pairs (model.cpp:145-205).
void Model::Store(const std::string& file) {
  std::unique_ptr<multiverso::Stream> s(
      multiverso::StreamFactory::GetStream(multiverso::URI(file), multiverso::FileOpenMode::BinaryWrite));
  if (!s || !s->Good()) {
    Log::Error("cannot write the model file %s\n", file.c_str());
    return;
  }
  if (!cfg_.sparse && !ftrl_) {
    s->Write(w_.data(), w_.size() * sizeof(float));
  } else {
    size_t count = 0;
    for (int64_t k = 0; k < size_; ++k) count += ftrl_ ? (z_[k] != 0 || n_[k] != 0) : (w_[k] != 0);
    s->Write(&count, sizeof count);
    for (int64_t k = 0; k < size_; ++k) {
      const size_t key = static_cast<size_t>(k);
      if (ftrl_) {
        if (z_[k] == 0 && n_[k] == 0) continue;
        const multiverso::FTRLEntry<float> e(z_[k], n_[k]);
        s->Write(&key, sizeof key);
        s->Write(&e, sizeof e);
      } else if (w_[k] != 0) {
        s->Write(&key, sizeof key);
        s->Write(&w_[k], sizeof(float));
      }
    }
  }
  s->Flush();
  Log::Info("model written to %s\n", file.c_str());
}

void Model::ReadModel(const std::string& file) {
  std::unique_ptr<multiverso::Stream> s(
      multiverso::StreamFactory::GetStream(multiverso::URI(file), multiverso::FileOpenMode::BinaryRead));
  if (!s || !s->Good()) Log::Fatal("cannot read the model file %s\n", file.c_str());
  if (!cfg_.sparse && !ftrl_) {
    if (s->Read(w_.data(), w_.size() * sizeof(float)) != w_.size() * sizeof(float))
      Log::Fatal("model file %s does not hold %lld weights\n", file.c_str(), static_cast<long long>(size_));
    return;
  }
  size_t count = 0;
  s->Read(&count, sizeof count);
  for (size_t i = 0; i < count; ++i) {
    size_t key = 0;
    s->Read(&key, sizeof key);
    if (key >= static_cast<size_t>(size_)) Log::Fatal("model file %s: key %zu out of range\n", file.c_str(), key);
    if (ftrl_) {
      multiverso::FTRLEntry<float> e;
      s->Read(&e, sizeof e);
      z_[key] = e.z;
      n_[key] = e.n;
      RefreshFtrlWeight(static_cast<int64_t>(key));
    } else {
      s->Read(&w_[key], sizeof(float));
    }
  }
}

void Model::Load(const std::string& file) {
  ReadModel(file);
  Log::Info("model loaded from %s\n", file.c_str());
}

void Model::LogTimes() const {
  if (updates_) Log::Info("average computation time: %.3fms per minibatch\n", compute_ms_ / updates_);
}

// ---------------------------------------------------------------------------------- PSModel
PSModel::PSModel(const Configure& config) : Model(config) {
  using namespace multiverso;
  if (ftrl_) {
    ftrl_table_.reset(MV_CreateTable(FTRLTableOption<float>(static_cast<size_t>(size_))));
  } else if (cfg_.sparse) {
    sparse_.reset(MV_CreateTable(SparseTableOption<float>(static_cast<size_t>(size_))));
  } else {
    dense_.reset(MV_CreateTable(ArrayTableOption<float>(static_cast<size_t>(size_))));
    next_w_.assign(size_, 0.0f);
  }
  if (dense_ == nullptr && sparse_ == nullptr && ftrl_table_ == nullptr)
    Log::Fatal("logreg: this rank is not a worker (run with the default -ps_role)\n");
}

void PSModel::ExpandKeys(const std::vector<int64_t>& feature_keys, std::vector<size_t>* table_keys) const {
  table_keys->clear();
  table_keys->reserve(feature_keys.size() * out_);
  for (int c = 0; c < out_; ++c)
    for (int64_t k : feature_keys) table_keys->push_back(static_cast<size_t>(c) * dim_ + k);
}

void PSModel::StartSparsePull(const std::vector<int64_t>& feature_keys, SparsePull* p) {
  ExpandKeys(feature_keys, &p->keys);
  if (ftrl_) {
    p->entries.assign(p->keys.size(), multiverso::FTRLEntry<float>());
    p->handle = ftrl_table_->GetAsync(p->keys.data(), p->keys.size(), p->entries.data());
  } else {
    p->vals.assign(p->keys.size(), 0.0f);
    p->handle = sparse_->GetAsync(p->keys.data(), p->keys.size(), p->vals.data());
  }
}

void PSModel::FinishSparsePull(SparsePull* p) {
  if (p->handle < 0) return;
  if (ftrl_) {
    ftrl_table_->Wait(p->handle);
    for (size_t i = 0; i < p->keys.size(); ++i) {
      z_[p->keys[i]] = p->entries[i].z;
      n_[p->keys[i]] = p->entries[i].n;
      RefreshFtrlWeight(static_cast<int64_t>(p->keys[i]));
    }
  } else {
    sparse_->Wait(p->handle);
    for (size_t i = 0; i < p->keys.size(); ++i) w_[p->keys[i]] = p->vals[i];
  }
  p->handle = -1;
}

void PSModel::BeginWindow(const std::vector<int64_t>& keys, const std::vector<int64_t>* next_keys) {
  Timer timer;
  ++windows_;
  if (dense_ != nullptr) {
    if (!cfg_.pipeline || dense_pending_ < 0) {
      dense_->Get(w_.data(), w_.size());
    } else {                               // GetPipelineTable: swap in the buffer requested last time
      dense_->Wait(dense_pending_);
      w_.swap(next_w_);
    }
    if (cfg_.pipeline) dense_pending_ = dense_->GetAsync(next_w_.data(), next_w_.size());
  } else {
    if (pending_.handle >= 0) {
      FinishSparsePull(&pending_);         // requested while the previous window trained
    } else {
      SparsePull now;
      StartSparsePull(keys, &now);
      FinishSparsePull(&now);
    }
    if (cfg_.pipeline && next_keys != nullptr && !next_keys->empty()) StartSparsePull(*next_keys, &pending_);
  }
  pull_ms_ += timer.elapse();
}

void PSModel::ApplyDelta() {
  Timer timer;
  Model::ApplyDelta();                     // keep the local copy moving between pulls
  if (dense_ != nullptr) {
    dense_->AddAsync(delta_.data(), delta_.size());
  } else {
    std::vector<size_t> keys(touched_.begin(), touched_.end());
    if (ftrl_) {
      std::vector<multiverso::FTRLEntry<float>> e(keys.size());
      for (size_t i = 0; i < keys.size(); ++i) e[i] = {delta_[keys[i]], delta_n_[keys[i]]};
      ftrl_table_->AddAsync(keys.data(), e.data(), keys.size());
    } else {
      std::vector<float> v(keys.size());
      for (size_t i = 0; i < keys.size(); ++i) v[i] = delta_[keys[i]];
      sparse_->AddAsync(keys.data(), v.data(), keys.size());
    }
  }
  push_ms_ += timer.elapse();
}

void PSModel::PullAll() {
  if (dense_ != nullptr) {
    if (dense_pending_ >= 0) {
      dense_->Wait(dense_pending_);
      dense_pending_ = -1;                 // the next window starts with a blocking Get again
    }
    dense_->Get(w_.data(), w_.size());
    return;
  }
  FinishSparsePull(&pending_);
  std::vector<size_t> keys;
  if (ftrl_) {
    std::vector<multiverso::FTRLEntry<float>> e;
    ftrl_table_->GetAll(&keys, &e);
    for (size_t i = 0; i < keys.size(); ++i) {
      z_[keys[i]] = e[i].z;
      n_[keys[i]] = e[i].n;
      RefreshFtrlWeight(static_cast<int64_t>(keys[i]));
    }
  } else {
    std::vector<float> v;
    sparse_->GetAll(&keys, &v);
    for (size_t i = 0; i < keys.size(); ++i) w_[keys[i]] = v[i];
  }
}

// PSModel::Load (ps_model.cpp:115-154): worker 0 pushes the file's model through the servers,
// negated because the server subtracts; then everyone pulls.
void PSModel::Load(const std::string& file) {
  PullAll();
  if (multiverso::MV_WorkerId() == 0) {
    std::vector<float> cur_w(w_), cur_z(z_), cur_n(n_);
    ReadModel(file);
    if (dense_ != nullptr) {
      std::vector<float> d(size_);
      for (int64_t k = 0; k < size_; ++k) d[k] = -(w_[k] - cur_w[k]);
      dense_->Add(d.data(), d.size());
    } else {
      std::vector<size_t> keys;
      std::vector<float> v;
      std::vector<multiverso::FTRLEntry<float>> e;
      for (int64_t k = 0; k < size_; ++k) {
        if (ftrl_) {
          if (z_[k] == cur_z[k] && n_[k] == cur_n[k]) continue;
          keys.push_back(k);
          e.emplace_back(-(z_[k] - cur_z[k]), -(n_[k] - cur_n[k]));
        } else if (w_[k] != cur_w[k]) {
          keys.push_back(k);
          v.push_back(-(w_[k] - cur_w[k]));
        }
      }
      if (!keys.empty()) {
        if (ftrl_) ftrl_table_->Add(keys.data(), e.data(), keys.size());
        else sparse_->Add(keys.data(), v.data(), keys.size());
      }
    }
    Log::Info("model %s pushed to the servers\n", file.c_str());
  }
  multiverso::MV_Barrier();
  PullAll();
}

void PSModel::Store(const std::string& file) {
  multiverso::MV_Barrier();
  PullAll();
  if (multiverso::MV_WorkerId() == 0) Model::Store(file);
}

void PSModel::LogTimes() const {
  Model::LogTimes();
  if (updates_)
    Log::Info("average communication time: push %.3fms per minibatch, pull %.3fms per window\n", push_ms_ / updates_,
              pull_ms_ / std::max<int64_t>(1, windows_));
}

}  // namespace logreg
