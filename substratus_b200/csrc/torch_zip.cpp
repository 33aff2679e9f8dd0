// torch_zip.cpp — reader for PyTorch `torch.save` checkpoints (`pytorch_model*.bin`), the other on-disk layout an HF
// snapshot can have under the Model artifact ("HF snapshot layout variants (*.bin vs safetensors,
// shard index)"; the reference's model-loader copies whatever the hub repo holds, examples/llama2-7b/base-model.yaml).
// Stands in for `torch.load` inside the external image's `from_pretrained`.  Format (torch >= 1.6):
//   * a ZIP archive (ZIP64 when > 4 GiB) whose members are STORED (no compression), 64-byte aligned:
//       <prefix>/data.pkl        pickle (protocol 2) of the state dict
//       <prefix>/data/<key>      raw little-endian bytes of storage <key>
//   * in the pickle every tensor is `torch._utils._rebuild_tensor_v2(storage, offset, size, stride, ...)` and every
//     storage a persistent id `('storage', torch.<T>Storage, key, location, numel)`.
// This file is a from-scratch reader of that container: a bounds-checked ZIP central-directory walk and a small
// pickle interpreter that knows only the opcodes/globals such files contain and REFUSES anything else (it never
// imports or calls anything — unlike torch.load, a hostile pickle can at worst be rejected).  No arithmetic here: the
// result is a set of TensorViews pointing into the mmap, consumed by the same upload path as safetensors.
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>

#include "loader.h"

namespace ssb {
namespace {

struct ZipEntry {
  const uint8_t* data = nullptr;
  uint64_t size = 0;
};

inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

// Central directory -> name -> (pointer, size) of every STORED member.
bool zip_index(const uint8_t* base, size_t n, std::map<std::string, ZipEntry>* out, std::string* err) {
  if (n < 22) return *err = "not a zip archive (too small)", false;
  // End Of Central Directory: last 22..22+65535 bytes
  size_t eocd = std::string::npos;
  const size_t lo = n > 22 + 65535 ? n - 22 - 65535 : 0;
  for (size_t p = n - 22 + 1; p-- > lo;)
    if (rd32(base + p) == 0x06054b50u) {
      eocd = p;
      break;
    }
  if (eocd == std::string::npos) return *err = "not a zip archive (legacy torch.save format is not supported; re-save with torch >= 1.6 or as safetensors)", false;
  uint64_t n_entries = rd16(base + eocd + 10), cd_size = rd32(base + eocd + 12), cd_off = rd32(base + eocd + 16);
  if (n_entries == 0xFFFF || cd_size == 0xFFFFFFFFu || cd_off == 0xFFFFFFFFu) {  // ZIP64: locator sits right before the EOCD
    if (eocd < 20 || rd32(base + eocd - 20) != 0x07064b50u) return *err = "zip64 locator missing", false;
    const uint64_t z64 = rd64(base + eocd - 20 + 8);
    if (z64 > n || n - z64 < 56 || rd32(base + z64) != 0x06064b50u) return *err = "zip64 end record missing", false;
    n_entries = rd64(base + z64 + 32);
    cd_size = rd64(base + z64 + 40);
    cd_off = rd64(base + z64 + 48);
  }
  if (cd_off > n || cd_size > n - cd_off) return *err = "zip central directory out of range", false;
  const uint8_t* p = base + cd_off;
  const uint8_t* end = p + cd_size;
  for (uint64_t i = 0; i < n_entries; ++i) {
    if (end - p < 46 || rd32(p) != 0x02014b50u) return *err = "corrupt zip central directory", false;
    const uint16_t method = rd16(p + 10), nlen = rd16(p + 28), elen = rd16(p + 30), clen = rd16(p + 32);
    uint64_t csize = rd32(p + 20), usize = rd32(p + 24), lho = rd32(p + 42);
    if ((size_t)(end - p) < 46u + nlen + elen + clen) return *err = "corrupt zip central directory", false;
    const std::string name((const char*)p + 46, nlen);
    const uint8_t* x = p + 46 + nlen;
    const uint8_t* xe = x + elen;
    while (xe - x >= 4) {  // ZIP64 extended information: the fields that overflowed, in this fixed order
      const uint16_t id = rd16(x), sz = rd16(x + 2);
      const uint8_t* f = x + 4;
      if (xe - f < sz) break;
      if (id == 0x0001) {
        const uint8_t* fe = f + sz;
        if (usize == 0xFFFFFFFFu && fe - f >= 8) usize = rd64(f), f += 8;
        if (csize == 0xFFFFFFFFu && fe - f >= 8) csize = rd64(f), f += 8;
        if (lho == 0xFFFFFFFFu && fe - f >= 8) lho = rd64(f), f += 8;
      }
      x += 4 + sz;
    }
    p += 46 + nlen + elen + clen;
    if (!name.empty() && name.back() == '/') continue;  // directory
    if (method != 0 || csize != usize) return *err = "zip member " + name + " is compressed (torch.save stores members uncompressed)", false;
    if (lho > n || n - lho < 30 || rd32(base + lho) != 0x04034b50u) return *err = "zip member " + name + ": bad local header", false;
    const uint64_t data_off = lho + 30 + rd16(base + lho + 26) + rd16(base + lho + 28);
    if (data_off > n || usize > n - data_off) return *err = "zip member " + name + " runs past the end of the file", false;
    (*out)[name] = ZipEntry{base + data_off, usize};
  }
  return true;
}

// ---- the values a state-dict pickle can hold
struct PVal;
using PRef = std::shared_ptr<PVal>;
struct PVal {
  enum Kind { None, Bool, Int, Float, Str, Seq, Dict, Global, Storage, Tensor, Opaque } kind = None;
  int64_t i = 0;   // Int / Bool; Storage: numel; Tensor: storage offset (elements)
  double f = 0;
  std::string s;   // Str; Global: "module.name"; Storage: key
  std::vector<PRef> seq;                      // Seq (tuple or list); Tensor: [storage]
  std::vector<std::pair<PRef, PRef>> dict;    // Dict
  int dtype = DT_OTHER;                       // Storage
  std::vector<int64_t> shape, stride;         // Tensor
};
PRef mk(PVal::Kind k) {
  auto v = std::make_shared<PVal>();
  v->kind = k;
  return v;
}

class Unpickler {
 public:
  Unpickler(const uint8_t* p, size_t n) : p_(p), e_(p + n) {}
  PRef run() {
    for (;;) {
      const uint8_t op = u8();
      switch (op) {
        case 0x80: u8(); break;                         // PROTO
        case 0x95: take(8); break;                      // FRAME
        case '.': return pop();                         // STOP
        case '(': marks_.push_back(st_.size()); break;  // MARK
        case 'N': st_.push_back(mk(PVal::None)); break;
        case 0x88: case 0x89: { auto v = mk(PVal::Bool); v->i = op == 0x88; st_.push_back(v); break; }
        case 'J': push_int((int32_t)rd32(take(4))); break;
        case 'K': push_int(u8()); break;
        case 'M': push_int(rd16(take(2))); break;
        case 0x8a: { const int n = u8(); push_int(long_le(take(n), n)); break; }   // LONG1
        case 0x8b: { const uint32_t n = rd32(take(4)); if (n > 8) fail("LONG4 wider than 64 bits"); push_int(long_le(take(n), (int)n)); break; }
        case 'G': { const uint8_t* b = take(8); uint64_t u = 0; for (int k = 0; k < 8; ++k) u = (u << 8) | b[k];
                    auto v = mk(PVal::Float); memcpy(&v->f, &u, 8); st_.push_back(v); break; }
        case 'X': push_str(rd32(take(4))); break;       // BINUNICODE
        case 0x8c: push_str(u8()); break;               // SHORT_BINUNICODE
        case 0x8d: push_str(rd64(take(8))); break;      // BINUNICODE8
        case 'U': case 'C': push_str(u8()); break;      // SHORT_BINSTRING / SHORT_BINBYTES
        case 'T': case 'B': push_str(rd32(take(4))); break;
        case 'c': { std::string m = line(), n = line(); auto v = mk(PVal::Global); v->s = m + "." + n; st_.push_back(v); break; }
        case 0x93: { PRef n = pop(), m = pop(); if (n->kind != PVal::Str || m->kind != PVal::Str) fail("STACK_GLOBAL on non-strings");
                     auto v = mk(PVal::Global); v->s = m->s + "." + n->s; st_.push_back(v); break; }
        case '}': st_.push_back(mk(PVal::Dict)); break;
        case ']': case ')': st_.push_back(mk(PVal::Seq)); break;
        case 't': { auto v = mk(PVal::Seq); v->seq = pop_mark(); st_.push_back(v); break; }
        case 0x85: case 0x86: case 0x87: { const size_t n = op - 0x84; if (st_.size() < n) fail("stack underflow");
                     auto v = mk(PVal::Seq); v->seq.assign(st_.end() - n, st_.end()); st_.resize(st_.size() - n); st_.push_back(v); break; }
        case 'q': memo_[u8()] = top(); break;
        case 'r': memo_[rd32(take(4))] = top(); break;
        case 0x94: { const uint64_t k = memo_.size(); memo_[k] = top(); break; }  // MEMOIZE
        case 'h': get(u8()); break;
        case 'j': get(rd32(take(4))); break;
        case 'a': { PRef x = pop(); if (top()->kind == PVal::Seq) top()->seq.push_back(x); break; }
        case 'e': {
          auto xs = pop_mark();
          if (top()->kind == PVal::Seq)
            for (auto& x : xs) top()->seq.push_back(x);
          break;
        }
        case 's': { PRef v = pop(), k = pop(); if (top()->kind == PVal::Dict) top()->dict.emplace_back(k, v); break; }
        case 'u': {
          auto xs = pop_mark();
          if (xs.size() % 2) fail("SETITEMS with an odd number of items");
          if (top()->kind == PVal::Dict)
            for (size_t k = 0; k < xs.size(); k += 2) top()->dict.emplace_back(xs[k], xs[k + 1]);
          break;
        }
        case 'b': pop(); (void)top(); break;            // BUILD: attribute state (OrderedDict._metadata, ...) is not needed
        case 0x81: { PRef args = pop(), cls = pop(); st_.push_back(construct(cls, args)); break; }  // NEWOBJ
        case 'R': { PRef args = pop(), fn = pop(); st_.push_back(construct(fn, args)); break; }     // REDUCE
        case 'Q': st_.push_back(storage(pop())); break;  // BINPERSID
        default: {
          char b[64];
          snprintf(b, sizeof b, "unsupported pickle opcode 0x%02x", op);
          fail(b);
        }
      }
    }
  }

 private:
  [[noreturn]] void fail(const std::string& m) { throw std::runtime_error("data.pkl: " + m); }
  const uint8_t* take(size_t n) {
    if ((size_t)(e_ - p_) < n) fail("truncated");
    const uint8_t* r = p_;
    p_ += n;
    return r;
  }
  uint8_t u8() { return *take(1); }
  std::string line() {
    const uint8_t* s = p_;
    while (p_ < e_ && *p_ != '\n') ++p_;
    if (p_ == e_) fail("truncated");
    return std::string((const char*)s, (const char*)p_++);
  }
  static int64_t long_le(const uint8_t* b, int n) {
    if (n == 0) return 0;
    if (n > 8) throw std::runtime_error("data.pkl: integer wider than 64 bits");
    uint64_t u = 0;
    for (int k = 0; k < n; ++k) u |= (uint64_t)b[k] << (8 * k);
    if (n < 8 && (b[n - 1] & 0x80)) u |= ~0ull << (8 * n);
    return (int64_t)u;
  }
  void push_int(int64_t x) { auto v = mk(PVal::Int); v->i = x; st_.push_back(v); }
  void push_str(uint64_t n) { auto v = mk(PVal::Str); const uint8_t* b = take((size_t)n); v->s.assign((const char*)b, (size_t)n); st_.push_back(v); }
  PRef pop() {
    if (st_.empty() || (!marks_.empty() && st_.size() <= marks_.back())) fail("stack underflow");
    PRef v = st_.back();
    st_.pop_back();
    return v;
  }
  PRef& top() {
    if (st_.empty()) fail("stack underflow");
    return st_.back();
  }
  std::vector<PRef> pop_mark() {
    if (marks_.empty() || marks_.back() > st_.size()) fail("no MARK");
    std::vector<PRef> xs(st_.begin() + marks_.back(), st_.end());
    st_.resize(marks_.back());
    marks_.pop_back();
    return xs;
  }
  void get(uint64_t k) {
    auto it = memo_.find(k);
    if (it == memo_.end()) fail("memo key not set");
    st_.push_back(it->second);
  }
  static std::vector<int64_t> ints(const PRef& t) {
    std::vector<int64_t> v;
    if (t->kind != PVal::Seq) throw std::runtime_error("data.pkl: expected a tuple of ints");
    for (auto& x : t->seq) {
      if (x->kind != PVal::Int) throw std::runtime_error("data.pkl: expected a tuple of ints");
      v.push_back(x->i);
    }
    return v;
  }
  // ('storage', torch.<T>Storage, key, location, numel)
  PRef storage(const PRef& pid) {
    if (pid->kind != PVal::Seq || pid->seq.size() < 5 || pid->seq[0]->kind != PVal::Str || pid->seq[0]->s != "storage" ||
        pid->seq[1]->kind != PVal::Global || pid->seq[2]->kind != PVal::Str || pid->seq[4]->kind != PVal::Int)
      fail("unexpected persistent id (not a torch storage)");
    auto v = mk(PVal::Storage);
    const std::string& t = pid->seq[1]->s;
    v->dtype = t == "torch.BFloat16Storage" ? DT_BF16 : t == "torch.HalfStorage" ? DT_F16 : t == "torch.FloatStorage" ? DT_F32 : DT_OTHER;
    v->s = pid->seq[2]->s;
    v->i = pid->seq[4]->i;
    return v;
  }
  PRef construct(const PRef& fn, const PRef& args) {
    if (fn->kind != PVal::Global || args->kind != PVal::Seq) return mk(PVal::Opaque);
    const std::string& g = fn->s;
    if (g == "collections.OrderedDict") {
      auto d = mk(PVal::Dict);
      if (!args->seq.empty() && args->seq[0]->kind == PVal::Seq)
        for (auto& kv : args->seq[0]->seq)
          if (kv->kind == PVal::Seq && kv->seq.size() == 2) d->dict.emplace_back(kv->seq[0], kv->seq[1]);
      return d;
    }
    if (g == "torch._utils._rebuild_tensor_v2" || g == "torch._utils._rebuild_tensor") {
      if (args->seq.size() < 4 || args->seq[0]->kind != PVal::Storage || args->seq[1]->kind != PVal::Int) fail("malformed tensor record");
      auto t = mk(PVal::Tensor);
      t->seq.push_back(args->seq[0]);
      t->i = args->seq[1]->i;
      t->shape = ints(args->seq[2]);
      t->stride = ints(args->seq[3]);
      return t;
    }
    if (g == "torch._utils._rebuild_parameter" && !args->seq.empty()) return args->seq[0];  // nn.Parameter wrapper
    return mk(PVal::Opaque);  // anything else (dtype objects, sizes, hooks ...) carries no weights
  }

  const uint8_t* p_;
  const uint8_t* e_;
  std::vector<PRef> st_;
  std::vector<size_t> marks_;
  std::map<uint64_t, PRef> memo_;
};

size_t elem_size(int dt) { return dt == DT_F32 ? 4 : 2; }

}  // namespace

bool ModelFiles::open_torch_zip(const std::string& path, std::string* err) {
  auto f = MappedFile::open(path, err);
  if (!f) return false;
  std::map<std::string, ZipEntry> members;
  std::string zerr;
  if (!zip_index(f->data(), f->size(), &members, &zerr)) return *err = path + ": " + zerr, false;
  std::string prefix;
  const ZipEntry* pkl = nullptr;
  for (auto& kv : members) {
    const std::string& n = kv.first;
    if (n == "data.pkl" || (n.size() > 9 && n.compare(n.size() - 9, 9, "/data.pkl") == 0)) {
      pkl = &kv.second;
      prefix = n.substr(0, n.size() - 8);
    }
  }
  if (!pkl) return *err = path + ": no data.pkl member (not a torch.save checkpoint)", false;
  auto bo = members.find(prefix + "byteorder");
  if (bo != members.end() && std::string((const char*)bo->second.data, (size_t)bo->second.size).rfind("little", 0) != 0)
    return *err = path + ": big-endian checkpoint", false;
  PRef root;
  try {
    root = Unpickler(pkl->data, (size_t)pkl->size).run();
  } catch (std::exception& e) {
    return *err = path + ": " + e.what(), false;
  }
  if (root->kind == PVal::Dict)  // {"state_dict": {...}} wrappers
    for (auto& kv : root->dict)
      if (kv.first->kind == PVal::Str && (kv.first->s == "state_dict" || kv.first->s == "model") && kv.second->kind == PVal::Dict) {
        root = kv.second;
        break;
      }
  if (root->kind != PVal::Dict) return *err = path + ": data.pkl does not hold a state dict", false;
  size_t added = 0;
  for (auto& kv : root->dict) {
    if (kv.first->kind != PVal::Str || kv.second->kind != PVal::Tensor) continue;  // _metadata, step counters, ...
    const PVal& t = *kv.second;
    const PVal& st = *t.seq[0];
    TensorView v;
    v.name = kv.first->s;
    v.dtype = st.dtype;
    v.shape = t.shape;
    int64_t numel = 1, expect = 1;
    bool contiguous = t.shape.size() == t.stride.size();
    for (size_t d = t.shape.size(); contiguous && d-- > 0;) {
      if (t.shape[d] < 0) return *err = path + ": " + v.name + ": negative dimension", false;
      if (t.shape[d] != 1 && t.stride[d] != expect) contiguous = false;
      expect *= t.shape[d];
    }
    for (int64_t s : t.shape) numel *= s;
    if (!contiguous) return *err = path + ": " + v.name + " is not stored contiguously", false;
    auto m = members.find(prefix + "data/" + st.s);
    if (m == members.end()) return *err = path + ": storage " + st.s + " of " + v.name + " is missing from the archive", false;
    if (st.dtype != DT_OTHER) {
      const size_t es = elem_size(st.dtype);
      if (t.i < 0 || (uint64_t)t.i * es > m->second.size || (uint64_t)numel * es > m->second.size - (uint64_t)t.i * es)
        return *err = path + ": " + v.name + " runs past its storage", false;
      v.data = m->second.data + (size_t)t.i * es;
      v.nbytes = (size_t)numel * es;
    }
    tensors_[v.name] = v;  // unsupported storage types stay listed (DT_OTHER) so the inventory check can name them
    ++added;
  }
  if (!added) return *err = path + ": no tensors in the state dict", false;
  files_.push_back(std::move(f));
  return true;
}

}  // namespace ssb
