// diskstore.hpp — the slice of diskstore.Bucket (diskstore/diskstore.go:45-65) the vector index
// touches, plus the in-memory bucket the reference's own tests use (diskstore/memstore.go:11-
// 113). The real bbolt-backed bucket stays in Go; over cgo the Go side hands the index the
// key/value pairs it needs, so this interface is what both sides agree on.
#pragma once
#include <functional>
#include <map>
#include <string>

namespace semadb {

// Go's `error`: empty message = nil.
struct Error {
  std::string msg;
  Error() = default;
  explicit Error(std::string m) : msg(std::move(m)) {}
  explicit operator bool() const { return !msg.empty(); }
  // fmt.Errorf("context: %w", err)
  Error Wrap(const std::string& context) const { return Error(context + ": " + msg); }
};
inline Error Ok() { return Error(); }

namespace diskstore {

using KVFn = std::function<Error(const std::string& k, const std::string& v)>;

class Bucket {
 public:
  virtual ~Bucket() = default;
  virtual bool IsReadOnly() const = 0;
  // nil slice in Go = false here
  virtual bool Get(const std::string& k, std::string* v) const = 0;
  virtual Error Put(const std::string& k, const std::string& v) = 0;
  virtual Error Delete(const std::string& k) = 0;
  virtual Error ForEach(const KVFn& f) const = 0;
  virtual Error PrefixScan(const std::string& prefix, const KVFn& f) const = 0;
};

// memstore.go:11-113 (ordered map here; Go's iteration order is unspecified anyway)
class MemBucket : public Bucket {
 public:
  explicit MemBucket(bool read_only = false) : ro_(read_only) {}
  bool IsReadOnly() const override { return ro_; }
  bool Get(const std::string& k, std::string* v) const override {
    auto it = data_.find(k);
    if (it == data_.end()) return false;
    if (v) *v = it->second;
    return true;
  }
  Error Put(const std::string& k, const std::string& v) override {
    if (ro_) return Error("cannot put into read-only memory bucket");
    data_[k] = v;
    return Ok();
  }
  Error Delete(const std::string& k) override {
    if (ro_) return Error("cannot delete in a read-only memory bucket");
    data_.erase(k);
    return Ok();
  }
  Error ForEach(const KVFn& f) const override {
    for (const auto& kv : data_)
      if (Error e = f(kv.first, kv.second)) return e;
    return Ok();
  }
  Error PrefixScan(const std::string& prefix, const KVFn& f) const override {
    for (auto it = data_.lower_bound(prefix); it != data_.end() && it->first.compare(0, prefix.size(), prefix) == 0; ++it)
      if (Error e = f(it->first, it->second)) return e;
    return Ok();
  }
  size_t Size() const { return data_.size(); }
  void SetReadOnly(bool ro) { ro_ = ro; }

 private:
  std::map<std::string, std::string> data_;
  bool ro_;
};

}  // namespace diskstore
}  // namespace semadb
