// tokenizer.h — native BPE tokenizer for the serve host's text path.
//
// The reference's only inference request carries TEXT (`POST /v1/completions {"prompt": "...", "max_tokens": 3}`,
// test/system.sh:73-78) and its external serving images tokenise with HF `tokenizers`.  This is a from-scratch reader
// of the same on-disk format (`tokenizer.json`, which the model-loader image stores next to the weights) for the two
// tokenizer families of the BASELINE configs:
//   * SentencePiece-style BPE with byte fallback (Llama-2: normalizer Prepend("▁") + Replace(" ", "▁"), no
//     pre-tokenizer, decoder Replace/ByteFallback/Fuse/Strip), and
//   * byte-level BPE with the GPT-2 split pattern (OPT-125m; ByteLevel pre-tokenizer + decoder), also inside the
//     pre-tokenizer Sequence Falcon ships (Punctuation(Contiguous), ByteLevel, Digits, Split([0-9][0-9][0-9])).
// Anything else (other normalizers / pre-tokenizers) is REFUSED at load time — the host then only accepts token ids —
// rather than tokenised approximately.  Parity oracle: the `tokenizers` library (tests/test_tokenizer.py).
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace ssb {

class Tokenizer {
 public:
  // Loads <dir>/tokenizer.json.  Returns false and sets *err if the file is missing or uses unsupported components.
  bool load(const std::string& path, std::string* err);
  // text -> ids (with the post-processor's special tokens, e.g. BOS, when add_special is true)
  std::vector<int32_t> encode(const std::string& text, bool add_special = true) const;
  // ids -> text (special tokens skipped when skip_special is true)
  std::string decode(const std::vector<int32_t>& ids, bool skip_special = true) const;
  int vocab_size() const { return (int)id_to_token_.size(); }
  bool byte_level() const { return byte_level_; }

 private:
  void bpe_word(const std::string& word, std::vector<int32_t>* out) const;
  void encode_segment(const std::string& text, bool first_segment, std::vector<int32_t>* out) const;
  static void gpt2_split(const std::string& text, std::vector<std::string>* pieces);
  struct PreStage {  // one pre-tokenizer of the byte-level family; each stage re-splits the previous stage's pieces
    enum Kind { kByteLevel, kPunct, kDigits, kSplitDigits3 } kind = kByteLevel;
    bool a = false;  // ByteLevel: add_prefix_space; Punctuation / Digits: contiguous (else isolated)
    bool b = false;  // ByteLevel: use_regex (the GPT-2 pattern)
  };
  std::vector<PreStage> pre_;

  std::unordered_map<std::string, int32_t> vocab_;
  std::vector<std::string> id_to_token_;
  std::unordered_map<uint64_t, std::pair<int32_t, int32_t>> merges_;  // (left id << 32 | right id) -> (rank, merged id)
  std::vector<std::pair<std::string, int32_t>> added_;                // added/special tokens matched verbatim in the text
  std::vector<bool> is_special_;
  std::vector<int32_t> pre_special_, post_special_;                   // TemplateProcessing "single": ids before / after $A
  bool byte_level_ = false;        // GPT-2 byte-level family
  bool sp_prepend_ = false;        // Llama family: Prepend("▁")
  bool byte_fallback_ = false;
  int32_t unk_id_ = -1;
  int32_t byte_tokens_[256];       // <0xXX> ids (byte fallback), -1 if absent
  std::string byte_to_unicode_[256];                       // GPT-2 bytes_to_unicode
  std::unordered_map<std::string, uint8_t> unicode_to_byte_;
};

}  // namespace ssb
