// ArrayTable<T> (see include/multiverso/table/array_table.h).
#include "multiverso/table/array_table.h"
#include "multiverso/multiverso.h"
#include "multiverso/util/log.h"
#include "multiverso/util/parallel_for.h"

namespace multiverso {

std::vector<size_t> RangeOffsets(size_t total, int num_servers) {
  std::vector<size_t> off(static_cast<size_t>(num_servers) + 1, 0);
  const size_t each = total / static_cast<size_t>(num_servers);
  if (each == 0) {
    // fewer elements than servers: one element per server for the first `total` servers
    for (int s = 0; s <= num_servers; ++s) off[s] = std::min<size_t>(s, total);
    return off;
  }
  for (int s = 0; s < num_servers; ++s) off[s] = each * s;
  off[num_servers] = total;   // the last server takes the remainder
  return off;
}

namespace {
const integer_t kWholeTable = -1;
}

template <typename T>
ArrayWorker<T>::ArrayWorker(size_t size)
    : size_(size), num_server_(MV_NumServers()), offsets_(RangeOffsets(size, MV_NumServers())) {
  CHECK(size_ >= 1);
  Log::Debug("worker %d created ArrayTable with %zu elements", MV_WorkerId(), size);
}

template <typename T>
int ArrayWorker<T>::GetAsync(T* data, size_t size) {
  CHECK(size == size_);
  const int id = NewRequest();
  {
    std::lock_guard<std::mutex> lk(dest_mu_);
    dest_[id] = data;
  }
  // re-use the generic submit path but with our pre-allocated id: emulate by direct message
  std::vector<Blob> blobs;
  blobs.emplace_back(&kWholeTable, sizeof(integer_t));
  MessagePtr msg(new Message());
  msg->set_src(MV_Rank());
  msg->set_type(MsgType::Request_Get);
  msg->set_msg_id(id);
  msg->set_table_id(table_id_);
  msg->data() = std::move(blobs);
  Zoo::Get()->SendTo("worker", msg);
  return id;
}

template <typename T>
void ArrayWorker<T>::Get(T* data, size_t size) { Wait(GetAsync(data, size)); }

template <typename T>
int ArrayWorker<T>::AddAsync(T* data, size_t size, const AddOption* option) {
  CHECK(size == size_);
  return WorkerTable::AddAsync(Blob(&kWholeTable, sizeof(integer_t)), Blob(data, size * sizeof(T)), option);
}

template <typename T>
void ArrayWorker<T>::Add(T* data, size_t size, const AddOption* option) { Wait(AddAsync(data, size, option)); }

template <typename T>
int ArrayWorker<T>::Partition(const std::vector<Blob>& kv, MsgType type,
                              std::unordered_map<int, std::vector<Blob>>* out) {
  const bool is_add = type == MsgType::Request_Add;
  CHECK(kv.size() >= (is_add ? 2u : 1u));
  const Blob* option = (kv.size() > (is_add ? 2u : 1u)) ? &kv.back() : nullptr;
  for (int s = 0; s < num_server_; ++s) {
    const size_t lo = offsets_[s], hi = offsets_[s + 1];
    if (hi == lo) continue;
    std::vector<Blob>& v = (*out)[s];
    v.push_back(kv[0]);
    if (is_add) v.emplace_back(kv[1].data() + lo * sizeof(T), (hi - lo) * sizeof(T));
    if (option) v.push_back(*option);
  }
  return static_cast<int>(out->size());
}

template <typename T>
void ArrayWorker<T>::ProcessReplyGet(std::vector<Blob>& reply, int msg_id) {
  CHECK(reply.size() == 3);
  const int sid = reply[2].As<int>(0);
  T* dst;
  {
    std::lock_guard<std::mutex> lk(dest_mu_);
    dst = dest_.at(msg_id);
  }
  CHECK(reply[1].size() == (offsets_[sid + 1] - offsets_[sid]) * sizeof(T));
  ParallelMemcpy(dst + offsets_[sid], reply[1].data(), reply[1].size());
}

template <typename T>
void ArrayWorker<T>::OnRequestDone(int msg_id) {
  std::lock_guard<std::mutex> lk(dest_mu_);
  dest_.erase(msg_id);
}

// ---------------------------------------------------------------------------------------
template <typename T>
ArrayServer<T>::ArrayServer(size_t size) : server_id_(MV_ServerId()) {
  auto off = RangeOffsets(size, MV_NumServers());
  storage_.assign(off[server_id_ + 1] - off[server_id_], T(0));
  updater_ = Updater<T>::GetUpdater(storage_.size());
  Log::Debug("server %d created ArrayTable shard with %zu of %zu elements", server_id_, storage_.size(), size);
}

template <typename T>
ArrayServer<T>::~ArrayServer() { delete updater_; }

template <typename T>
void ArrayServer<T>::ProcessAdd(const std::vector<Blob>& data) {
  CHECK(data.size() >= 2);
  CHECK(data[1].size() == storage_.size() * sizeof(T));
  AddOption opt = AddOptionFrom(data, 2);
  updater_->Update(storage_.size(), storage_.data(), reinterpret_cast<T*>(data[1].data()), &opt, 0);
}

template <typename T>
void ArrayServer<T>::ProcessGet(const std::vector<Blob>& data, std::vector<Blob>* result) {
  CHECK(!data.empty());
  Blob values(storage_.size() * sizeof(T));
  updater_->Access(storage_.size(), storage_.data(), reinterpret_cast<T*>(values.data()), 0, nullptr);
  result->push_back(data[0]);
  result->push_back(std::move(values));
  result->emplace_back(&server_id_, sizeof(int));
}

template <typename T>
void ArrayServer<T>::Store(Stream* s) {
  s->Write(storage_.data(), storage_.size() * sizeof(T));
  std::vector<char> st(updater_->StateBytes());
  if (!st.empty()) {
    updater_->StoreState(st.data());
    s->Write(st.data(), st.size());
  }
}

template <typename T>
void ArrayServer<T>::Load(Stream* s) {
  // a truncated checkpoint must not leave the shard half loaded and report success: read into a scratch
  // buffer, install only a complete shard (reference: raw dump, no length check, array_table.cpp:147-151)
  std::vector<T> shard(storage_.size());
  const size_t want = shard.size() * sizeof(T);
  if (s->Read(shard.data(), want) != want) {
    Log::Error("table checkpoint is shorter than the shard (%zu bytes expected): not loaded\n", want);
    s->MarkFailed();
    return;
  }
  std::copy(shard.begin(), shard.end(), storage_.begin());
  std::vector<char> st(updater_->StateBytes());
  if (!st.empty()) {
    if (s->Read(st.data(), st.size()) == st.size()) updater_->LoadState(st.data());
    else Log::Info("table checkpoint carries no updater state (reference-format file): state left as is\n");
  }
}

template class ArrayWorker<float>;
template class ArrayWorker<double>;
template class ArrayWorker<int>;
template class ArrayServer<float>;
template class ArrayServer<double>;
template class ArrayServer<int>;

}  // namespace multiverso
