// Message: int header[8] = {src, dst, type, table_id, msg_id, ...} + vector<Blob>
// (include/multiverso/message.h:13-68). Replies negate the type and swap src/dst.
#ifndef MULTIVERSO_MESSAGE_H_
#define MULTIVERSO_MESSAGE_H_
#include <memory>
#include <vector>
#include "multiverso/blob.h"

namespace multiverso {

enum class MsgType : int {
  Request_Get = 1,
  Request_Add = 2,
  Reply_Get = -1,
  Reply_Add = -2,
  Server_Finish_Train = 31,
  Control_Barrier = 33,
  Control_Reply_Barrier = -33,
  Control_Register = 34,
  Control_Reply_Register = -34,
  Default = 0
};

class Message {
 public:
  static constexpr int kHeaderSize = 8;
  Message() { for (int& h : header_) h = 0; }
  int src() const { return header_[0]; }
  int dst() const { return header_[1]; }
  MsgType type() const { return static_cast<MsgType>(header_[2]); }
  int table_id() const { return header_[3]; }
  int msg_id() const { return header_[4]; }
  void set_src(int v) { header_[0] = v; }
  void set_dst(int v) { header_[1] = v; }
  void set_type(MsgType t) { header_[2] = static_cast<int>(t); }
  void set_table_id(int v) { header_[3] = v; }
  void set_msg_id(int v) { header_[4] = v; }
  int* header() { return header_; }
  const int* header() const { return header_; }
  std::vector<Blob>& data() { return data_; }
  const std::vector<Blob>& data() const { return data_; }
  size_t size() const { return data_.size(); }
  void Push(const Blob& b) { data_.push_back(b); }
  void Push(Blob&& b) { data_.push_back(std::move(b)); }
  // header-only reply: swapped endpoints, negated type, same table / msg id
  Message* CreateReplyMessage() const {
    Message* r = new Message();
    r->set_src(dst());
    r->set_dst(src());
    r->header_[2] = -header_[2];
    r->set_table_id(table_id());
    r->set_msg_id(msg_id());
    return r;
  }

 private:
  int header_[kHeaderSize];
  std::vector<Blob> data_;
};

using MessagePtr = std::unique_ptr<Message>;

}  // namespace multiverso
#endif
