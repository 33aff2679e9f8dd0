// loader.h — model-directory readers: HF snapshot (config.json + safetensors or pytorch_model*.bin [+index]) and GGUF.
// Stands in for the external image's `from_pretrained(/content/model)`; the on-disk layouts are what the
// reference's model-loader image leaves under the Model artifact (SURVEY.md §8a D0, §8f #2;
// examples/llama2-7b/base-model.yaml, examples/llama2-13b-chat-gguf/base-model.yaml:8-9 `files: model.bin`).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "json.h"

namespace ssb {

enum DType : int { DT_BF16 = 0, DT_F16 = 1, DT_F32 = 2, DT_Q4_0 = 10, DT_Q4_K = 11, DT_Q6_K = 12, DT_Q8_0 = 13, DT_OTHER = 99 };

struct TensorView {
  std::string name;
  int dtype = DT_OTHER;
  std::vector<int64_t> shape;  // row-major, outermost first (GGUF dims are reversed into this order)
  const uint8_t* data = nullptr;
  size_t nbytes = 0;
  int64_t rows() const { return shape.size() >= 2 ? shape[0] : 1; }
  int64_t cols() const { return shape.empty() ? 0 : shape.back(); }
};

class MappedFile {
 public:
  ~MappedFile();
  static std::unique_ptr<MappedFile> open(const std::string& path, std::string* err);
  const uint8_t* data() const { return data_; }
  size_t size() const { return size_; }

 private:
  const uint8_t* data_ = nullptr;
  size_t size_ = 0;
  int fd_ = -1;
};

class ModelFiles {
 public:
  // Opens every *.safetensors — else every pytorch_model*.bin (torch.save zip, torch_zip.cpp) — else the single
  // *.gguf / model.bin GGUF under dir.  Returns false + err on failure.
  bool open(const std::string& dir, std::string* err);
  // One safetensors file (a tensor-parallel rank's pre-sharded artifact, <dir>/ssb_tp<N>/rank<r>.safetensors).
  bool open_file(const std::string& path, std::string* err) { return open_safetensors(path, err); }
  // string entries of the safetensors "__metadata__" objects seen so far (later files override earlier ones)
  const std::map<std::string, std::string>& metadata() const { return metadata_; }
  const TensorView* find(const std::string& name) const;
  bool is_gguf() const { return is_gguf_; }
  const Json& gguf_meta() const { return gguf_meta_; }  // GGUF key/values as a JSON object (numbers/strings/arrays)
  size_t n_tensors() const { return tensors_.size(); }
  const std::map<std::string, TensorView>& tensors() const { return tensors_; }

 private:
  bool open_safetensors(const std::string& path, std::string* err);
  bool open_gguf(const std::string& path, std::string* err);
  bool open_torch_zip(const std::string& path, std::string* err);
  std::vector<std::unique_ptr<MappedFile>> files_;
  std::map<std::string, TensorView> tensors_;
  std::map<std::string, std::string> metadata_;
  bool is_gguf_ = false;
  Json gguf_meta_;
};

bool read_text_file(const std::string& path, std::string* out);

}  // namespace ssb
