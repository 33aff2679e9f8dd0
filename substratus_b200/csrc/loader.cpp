// loader.cpp — mmap readers for safetensors and GGUF containers (formats only; no arithmetic here).
#include "loader.h"

#include <dirent.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <fstream>
#include <sstream>

namespace ssb {

MappedFile::~MappedFile() {
  if (data_) munmap(const_cast<uint8_t*>(data_), size_);
  if (fd_ >= 0) close(fd_);
}

std::unique_ptr<MappedFile> MappedFile::open(const std::string& path, std::string* err) {
  int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) {
    *err = "cannot open " + path + ": " + strerror(errno);
    return nullptr;
  }
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size == 0) {
    *err = "cannot stat (or empty) " + path;
    close(fd);
    return nullptr;
  }
  void* p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  if (p == MAP_FAILED) {
    *err = "mmap failed for " + path;
    close(fd);
    return nullptr;
  }
  madvise(p, (size_t)st.st_size, MADV_SEQUENTIAL);
  std::unique_ptr<MappedFile> f(new MappedFile());
  f->data_ = (const uint8_t*)p;
  f->size_ = (size_t)st.st_size;
  f->fd_ = fd;
  return f;
}

bool read_text_file(const std::string& path, std::string* out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::stringstream ss;
  ss << f.rdbuf();
  *out = ss.str();
  return true;
}

static bool ends_with(const std::string& s, const std::string& suf) {
  return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

bool ModelFiles::open(const std::string& dir, std::string* err) {
  std::vector<std::string> st, gg, pt;
  DIR* d = opendir(dir.c_str());
  if (!d) {
    *err = "cannot open model dir " + dir + ": " + strerror(errno);
    return false;
  }
  while (dirent* e = readdir(d)) {
    std::string n = e->d_name;
    if (ends_with(n, ".safetensors")) st.push_back(n);
    if (ends_with(n, ".gguf") || n == "model.bin") gg.push_back(n);
    if (n.rfind("pytorch_model", 0) == 0 && ends_with(n, ".bin")) pt.push_back(n);  // pytorch_model.bin / -0000x-of-0000y.bin
  }
  closedir(d);
  std::sort(st.begin(), st.end());
  std::sort(gg.begin(), gg.end());
  std::sort(pt.begin(), pt.end());
  if (!st.empty()) {
    for (auto& n : st)
      if (!open_safetensors(dir + "/" + n, err)) return false;
    return true;
  }
  if (!pt.empty()) {  // same tensor names as the safetensors of the same snapshot (HF writes both from one state dict)
    for (auto& n : pt)
      if (!open_torch_zip(dir + "/" + n, err)) return false;
    return true;
  }
  for (auto& n : gg) {
    auto f = MappedFile::open(dir + "/" + n, err);
    if (!f) return false;
    bool magic = f->size() >= 4 && memcmp(f->data(), "GGUF", 4) == 0;
    f.reset();
    if (magic) return open_gguf(dir + "/" + n, err);
  }
  *err = "no *.safetensors, pytorch_model*.bin or GGUF file in " + dir;
  return false;
}

const TensorView* ModelFiles::find(const std::string& name) const {
  auto it = tensors_.find(name);
  return it == tensors_.end() ? nullptr : &it->second;
}

bool ModelFiles::open_safetensors(const std::string& path, std::string* err) {
  auto f = MappedFile::open(path, err);
  if (!f) return false;
  if (f->size() < 8) {
    *err = path + ": truncated";
    return false;
  }
  uint64_t hlen;
  memcpy(&hlen, f->data(), 8);
  if (hlen > f->size() - 8) {
    *err = path + ": bad header length";
    return false;
  }
  Json hdr;
  try {
    hdr = JsonParser((const char*)f->data() + 8, (size_t)hlen).parse();
  } catch (std::exception& e) {
    *err = path + ": " + e.what();
    return false;
  }
  const uint8_t* base = f->data() + 8 + hlen;
  const size_t avail = f->size() - 8 - (size_t)hlen;
  for (auto& kv : hdr.obj) {
    if (kv.first == "__metadata__") {
      if (kv.second.kind == Json::Obj)
        for (auto& m : kv.second.obj)
          if (m.second.kind == Json::Str) metadata_[m.first] = m.second.str;
      continue;
    }
    const Json& t = kv.second;
    TensorView v;
    v.name = kv.first;
    std::string dt = t.get_str("dtype", "");
    v.dtype = dt == "BF16" ? DT_BF16 : dt == "F16" ? DT_F16 : dt == "F32" ? DT_F32 : DT_OTHER;
    const Json* sh = t.find("shape");
    const Json* off = t.find("data_offsets");
    if (!sh || !off || off->arr.size() != 2) {
      *err = path + ": malformed entry " + kv.first;
      return false;
    }
    if (sh->kind != Json::Arr) {
      *err = path + ": shape of " + kv.first + " is not an array";
      return false;
    }
    uint64_t numel = 1;
    for (auto& d : sh->arr) {
      if (d.kind != Json::Num || d.num < 0 || d.num != (double)(int64_t)d.num || (d.num > 0 && numel > (1ull << 48) / (uint64_t)d.num)) {
        *err = path + ": bad dimension in the shape of " + kv.first;
        return false;
      }
      v.shape.push_back((int64_t)d.num);
      numel *= (uint64_t)d.num;
    }
    uint64_t b = (uint64_t)off->arr[0].num, e = (uint64_t)off->arr[1].num;
    if (e < b || e > avail) {
      *err = path + ": data_offsets out of range for " + kv.first;
      return false;
    }
    // a truncated or corrupt shard must fail here (SSB_EIO), not as a device-side read past the staging buffer
    const uint64_t esz = v.dtype == DT_F32 ? 4 : (v.dtype == DT_BF16 || v.dtype == DT_F16) ? 2 : 0;
    if (esz && numel * esz != e - b) {
      *err = path + ": " + kv.first + " stores " + std::to_string(e - b) + " bytes but its shape and dtype need " + std::to_string(numel * esz);
      return false;
    }
    v.data = base + b;
    v.nbytes = (size_t)(e - b);
    tensors_[v.name] = v;
  }
  files_.push_back(std::move(f));
  return true;
}

// ---------------------------------------------------------------- GGUF (v2/v3), llama.cpp's container
namespace {
struct Cur {
  const uint8_t* p;
  const uint8_t* e;
  bool ok = true;
  template <typename T>
  T rd() {
    T v{};
    if ((size_t)(e - p) < sizeof(T)) {
      ok = false;
      return v;
    }
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  std::string str() {
    uint64_t n = rd<uint64_t>();
    if (!ok || (uint64_t)(e - p) < n) {
      ok = false;
      return "";
    }
    std::string s((const char*)p, (size_t)n);
    p += n;
    return s;
  }
};

Json gguf_value(Cur& c, uint32_t type, int depth = 0) {
  Json j;
  j.kind = Json::Num;
  switch (type) {
    case 0: j.num = c.rd<uint8_t>(); break;
    case 1: j.num = c.rd<int8_t>(); break;
    case 2: j.num = c.rd<uint16_t>(); break;
    case 3: j.num = c.rd<int16_t>(); break;
    case 4: j.num = c.rd<uint32_t>(); break;
    case 5: j.num = c.rd<int32_t>(); break;
    case 6: j.num = c.rd<float>(); break;
    case 7:
      j.kind = Json::Bool;
      j.b = c.rd<uint8_t>() != 0;
      break;
    case 8:
      j.kind = Json::Str;
      j.str = c.str();
      break;
    case 9: {
      uint32_t et = c.rd<uint32_t>();
      uint64_t n = c.rd<uint64_t>();
      j.kind = Json::Arr;
      // vocab arrays can hold 10^5 strings; keep them (the tokenizer host reads them) but guard depth
      if (depth > 2) {
        c.ok = false;
        break;
      }
      j.arr.reserve((size_t)std::min<uint64_t>(n, 1u << 20));
      for (uint64_t i = 0; i < n && c.ok; ++i) j.arr.push_back(gguf_value(c, et, depth + 1));
      break;
    }
    case 10: j.num = (double)c.rd<uint64_t>(); break;
    case 11: j.num = (double)c.rd<int64_t>(); break;
    case 12: j.num = c.rd<double>(); break;
    default: c.ok = false;
  }
  return j;
}
}  // namespace

bool ModelFiles::open_gguf(const std::string& path, std::string* err) {
  auto f = MappedFile::open(path, err);
  if (!f) return false;
  Cur c{f->data(), f->data() + f->size()};
  uint32_t magic = c.rd<uint32_t>();
  uint32_t version = c.rd<uint32_t>();
  (void)magic;
  if (version < 2 || version > 3) {
    *err = path + ": unsupported GGUF version " + std::to_string(version);
    return false;
  }
  uint64_t n_tensors = c.rd<uint64_t>();
  uint64_t n_kv = c.rd<uint64_t>();
  gguf_meta_.kind = Json::Obj;
  for (uint64_t i = 0; i < n_kv && c.ok; ++i) {
    std::string k = c.str();
    uint32_t t = c.rd<uint32_t>();
    gguf_meta_.obj.emplace_back(k, gguf_value(c, t));
  }
  struct Info {
    std::string name;
    std::vector<int64_t> dims;
    uint32_t type;
    uint64_t off;
  };
  std::vector<Info> infos;
  for (uint64_t i = 0; i < n_tensors && c.ok; ++i) {
    Info in;
    in.name = c.str();
    uint32_t nd = c.rd<uint32_t>();
    if (nd > 8) {  // GGML_MAX_DIMS is 4: a larger count is a corrupt header, not a reason to loop 2^32 times
      c.ok = false;
      break;
    }
    for (uint32_t d = 0; d < nd && c.ok; ++d) in.dims.push_back((int64_t)c.rd<uint64_t>());
    in.type = c.rd<uint32_t>();
    in.off = c.rd<uint64_t>();
    infos.push_back(std::move(in));
  }
  if (!c.ok) {
    *err = path + ": truncated GGUF header";
    return false;
  }
  const double align_d = gguf_meta_.get_num("general.alignment", 32);
  if (!(align_d >= 1 && align_d <= 65536)) {
    *err = path + ": bad general.alignment";
    return false;
  }
  const uint64_t align = (uint64_t)align_d;
  uint64_t pos = (uint64_t)(c.p - f->data());
  pos = (pos + align - 1) / align * align;
  if (pos > f->size()) {
    *err = path + ": truncated GGUF header";
    return false;
  }
  const uint8_t* base = f->data() + pos;
  const uint64_t avail = f->size() - pos;
  for (auto& in : infos) {
    TensorView v;
    v.name = in.name;
    v.shape.assign(in.dims.rbegin(), in.dims.rend());  // GGUF stores the fastest dim first
    int64_t n = 1;
    bool sane = true;
    for (auto d : in.dims) {  // corrupt dims must not overflow the byte count below
      if (d < 0 || d > (int64_t)1 << 40 || (d > 0 && n > ((int64_t)1 << 48) / d)) sane = false;
      if (sane) n *= d;
    }
    if (!sane) {
      *err = path + ": tensor " + in.name + " has impossible dimensions";
      return false;
    }
    switch (in.type) {
      case 0: v.dtype = DT_F32; v.nbytes = (size_t)n * 4; break;
      case 1: v.dtype = DT_F16; v.nbytes = (size_t)n * 2; break;
      case 30: v.dtype = DT_BF16; v.nbytes = (size_t)n * 2; break;
      case 2: v.dtype = DT_Q4_0; v.nbytes = (size_t)(n / 32) * 18; break;
      case 8: v.dtype = DT_Q8_0; v.nbytes = (size_t)(n / 32) * 34; break;
      case 12: v.dtype = DT_Q4_K; v.nbytes = (size_t)(n / 256) * 144; break;
      case 14: v.dtype = DT_Q6_K; v.nbytes = (size_t)(n / 256) * 210; break;
      default: v.dtype = DT_OTHER; v.nbytes = 0;
    }
    if (in.off > avail || v.nbytes > avail - in.off) {
      *err = path + ": tensor " + in.name + " out of range";
      return false;
    }
    v.data = base + in.off;
    tensors_[v.name] = v;
  }
  is_gguf_ = true;
  files_.push_back(std::move(f));
  return true;
}

}  // namespace ssb
