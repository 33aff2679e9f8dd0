// tokenizer.cpp — see tokenizer.h.  Algorithms restated from the HF `tokenizers` BPE model (merge the lowest-rank
// adjacent pair, leftmost first, until no pair is in the merge table), its ByteFallback / ByteLevel components and the
// GPT-2 split pattern  's|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+ .
#include "tokenizer.h"

#include <algorithm>
#include <cstring>
#include <queue>

#include "json.h"
#include "loader.h"
#include "unicode_tables.h"

namespace ssb {
namespace {

// ---- UTF-8 helpers
size_t utf8_len(unsigned char c) { return c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 1; }

uint32_t utf8_decode(const std::string& s, size_t i, size_t* n) {
  const unsigned char c = (unsigned char)s[i];
  size_t len = utf8_len(c);
  if (i + len > s.size()) len = 1;
  *n = len;
  if (len == 1) return c;
  uint32_t cp = c & (0xFF >> (len + 1));
  for (size_t k = 1; k < len; ++k) {
    const unsigned char d = (unsigned char)s[i + k];
    if ((d & 0xC0) != 0x80) {
      *n = 1;
      return c;
    }
    cp = (cp << 6) | (d & 0x3F);
  }
  return cp;
}

void utf8_append(std::string& o, uint32_t cp) {
  if (cp < 0x80)
    o += (char)cp;
  else if (cp < 0x800) {
    o += (char)(0xC0 | (cp >> 6));
    o += (char)(0x80 | (cp & 0x3F));
  } else if (cp < 0x10000) {
    o += (char)(0xE0 | (cp >> 12));
    o += (char)(0x80 | ((cp >> 6) & 0x3F));
    o += (char)(0x80 | (cp & 0x3F));
  } else {
    o += (char)(0xF0 | (cp >> 18));
    o += (char)(0x80 | ((cp >> 12) & 0x3F));
    o += (char)(0x80 | ((cp >> 6) & 0x3F));
    o += (char)(0x80 | (cp & 0x3F));
  }
}

// bytes -> string, every byte of an ill-formed sequence replaced by U+FFFD (String::from_utf8_lossy granularity is
// per maximal invalid prefix; the ByteFallback decoder of `tokenizers` emits one U+FFFD per byte token — callers pick)
bool utf8_valid(const std::string& s) {
  size_t i = 0;
  while (i < s.size()) {
    const unsigned char c = (unsigned char)s[i];
    size_t len = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 0;
    if (len == 0 || i + len > s.size()) return false;
    for (size_t k = 1; k < len; ++k)
      if (((unsigned char)s[i + k] & 0xC0) != 0x80) return false;
    if (len == 2 && c < 0xC2) return false;
    if (len == 3) {
      const uint32_t cp = ((c & 0x0F) << 12) | (((unsigned char)s[i + 1] & 0x3F) << 6) | ((unsigned char)s[i + 2] & 0x3F);
      if (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF)) return false;
    }
    if (len == 4) {
      const uint32_t cp = ((c & 0x07) << 18) | (((unsigned char)s[i + 1] & 0x3F) << 12) | (((unsigned char)s[i + 2] & 0x3F) << 6) |
                          ((unsigned char)s[i + 3] & 0x3F);
      if (cp < 0x10000 || cp > 0x10FFFF) return false;
    }
    i += len;
  }
  return true;
}

// ---- Unicode classes: \\p{L} / \\p{N} / \\s of the GPT-2 pattern, `Punctuation`'s and `Digits`' predicates.  The tables
// are generated from the behaviour of the HF `tokenizers` library itself (tools/gen_unicode_tables.py), all planes.
bool in_table(const uint32_t (*t)[2], int n, uint32_t c) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) / 2;
    if (c < t[mid][0])
      hi = mid - 1;
    else if (c > t[mid][1])
      lo = mid + 1;
    else
      return true;
  }
  return false;
}
bool is_space(uint32_t c) { return c < 0x80 ? (c == ' ' || (c >= 9 && c <= 13)) : in_table(utab::kSpace, utab::kSpaceN, c); }
bool is_number(uint32_t c) { return c < 0x80 ? (c >= '0' && c <= '9') : in_table(utab::kNumber, utab::kNumberN, c); }
bool is_letter(uint32_t c) {
  return c < 0x80 ? ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z')) : in_table(utab::kLetter, utab::kLetterN, c);
}
bool is_punct(uint32_t c) { return in_table(utab::kPunct, utab::kPunctN, c); }
bool is_numeric(uint32_t c) { return c < 0x80 ? (c >= '0' && c <= '9') : in_table(utab::kNumeric, utab::kNumericN, c); }

// split `s` at the characters `pred` selects: every selected char alone (isolated) or runs of them merged (contiguous);
// the text between is kept; no empty pieces  (tokenizers: NormalizedString::split with SplitDelimiterBehavior)
template <class Pred>
void split_chars(const std::string& s, Pred pred, bool contiguous, std::vector<std::string>* out) {
  size_t start = 0;   // start of the pending non-matching text
  size_t run = std::string::npos;  // start of the pending run of matches (contiguous only)
  for (size_t i = 0; i < s.size();) {
    size_t n;
    const uint32_t cp = utf8_decode(s, i, &n);
    if (pred(cp)) {
      if (run == std::string::npos) {
        if (i > start) out->push_back(s.substr(start, i - start));
        run = i;
      }
      if (!contiguous) {
        out->push_back(s.substr(i, n));
        run = std::string::npos;
      }
      start = i + n;
    } else if (run != std::string::npos) {
      if (contiguous) out->push_back(s.substr(run, i - run));
      run = std::string::npos;
      start = i;
    }
    i += n;
  }
  if (run != std::string::npos && contiguous) out->push_back(s.substr(run));
  else if (start < s.size()) out->push_back(s.substr(start));
}

struct Sym {
  int32_t id;
  int prev, next;
};
struct Cand {
  int32_t rank;
  int pos;
  int32_t left, right;
  bool operator>(const Cand& o) const { return rank != o.rank ? rank > o.rank : pos > o.pos; }
};

}  // namespace

bool Tokenizer::load(const std::string& path, std::string* err) {
  std::string txt;
  if (!read_text_file(path, &txt)) {
    *err = "cannot read " + path;
    return false;
  }
  Json j;
  try {
    j = json_parse(txt);
  } catch (std::exception& e) {
    *err = path + ": " + e.what();
    return false;
  }
  const Json* model = j.find("model");
  if (!model || model->get_str("type", "BPE") != "BPE") {
    *err = "tokenizer model is not BPE";
    return false;
  }
  if (model->has("continuing_subword_prefix") || model->has("end_of_word_suffix") || model->has("dropout")) {
    *err = "BPE options continuing_subword_prefix / end_of_word_suffix / dropout are not supported";
    return false;
  }
  const Json* vocab = model->find("vocab");
  const Json* merges = model->find("merges");
  if (!vocab || vocab->kind != Json::Obj || !merges || merges->kind != Json::Arr) {
    *err = "tokenizer.json: missing vocab/merges";
    return false;
  }
  size_t max_id = 0;
  for (auto& kv : vocab->obj) max_id = std::max(max_id, (size_t)kv.second.num);
  id_to_token_.assign(max_id + 1, "");
  for (auto& kv : vocab->obj) {
    vocab_[kv.first] = (int32_t)kv.second.num;
    id_to_token_[(size_t)kv.second.num] = kv.first;
  }
  int32_t rank = 0;
  for (auto& m : merges->arr) {
    std::string a, b;
    if (m.kind == Json::Arr && m.arr.size() == 2) {
      a = m.arr[0].str;
      b = m.arr[1].str;
    } else if (m.kind == Json::Str) {
      const size_t sp = m.str.find(' ');
      if (sp == std::string::npos) continue;
      a = m.str.substr(0, sp);
      b = m.str.substr(sp + 1);
    } else {
      continue;
    }
    auto ia = vocab_.find(a), ib = vocab_.find(b), ic = vocab_.find(a + b);
    if (ia != vocab_.end() && ib != vocab_.end() && ic != vocab_.end())
      merges_[((uint64_t)(uint32_t)ia->second << 32) | (uint32_t)ib->second] = {rank, ic->second};
    ++rank;
  }
  byte_fallback_ = model->get_num("byte_fallback", 0) != 0;
  const std::string unk = model->get_str("unk_token", "");
  if (!unk.empty() && vocab_.count(unk)) unk_id_ = vocab_[unk];
  for (int b = 0; b < 256; ++b) {
    char name[8];
    snprintf(name, sizeof name, "<0x%02X>", b);
    auto it = vocab_.find(name);
    byte_tokens_[b] = it == vocab_.end() ? -1 : it->second;
  }
  // added tokens
  is_special_.assign(id_to_token_.size(), false);
  if (const Json* at = j.find("added_tokens"))
    for (auto& t : at->arr) {
      const int32_t id = (int32_t)t.get_int("id", -1);
      const std::string content = t.get_str("content", "");
      if (id < 0 || content.empty()) continue;
      if ((size_t)id >= id_to_token_.size()) {
        id_to_token_.resize(id + 1);
        is_special_.resize(id + 1, false);
      }
      id_to_token_[id] = content;
      vocab_[content] = id;
      added_.emplace_back(content, id);
      if (t.get_num("special", 0) != 0) is_special_[id] = true;
    }
  std::sort(added_.begin(), added_.end(), [](auto& a, auto& b) { return a.first.size() > b.first.size(); });
  // normalizer
  const Json* norm = j.find("normalizer");
  const Json* pre = j.find("pre_tokenizer");
  if (norm && !norm->is_null()) {
    // accepted: Sequence[Prepend("▁"), Replace(" ", "▁")] (Llama-2) or those two alone
    std::vector<const Json*> items;
    if (norm->get_str("type", "") == "Sequence") {
      if (const Json* ns = norm->find("normalizers"))
        for (auto& n : ns->arr) items.push_back(&n);
    } else {
      items.push_back(norm);
    }
    bool have_replace = false;
    for (const Json* n : items) {
      const std::string t = n->get_str("type", "");
      if (t == "Prepend" && n->get_str("prepend", "") == "\xE2\x96\x81") {
        sp_prepend_ = true;
      } else if (t == "Replace" && n->get_str("content", "") == "\xE2\x96\x81" && n->find("pattern") &&
                 n->find("pattern")->get_str("String", "") == " ") {
        have_replace = true;
      } else {
        *err = "unsupported normalizer '" + t + "'";
        return false;
      }
    }
    if (!have_replace) {
      *err = "unsupported normalizer sequence (expected Replace(' ', U+2581))";
      return false;
    }
    if (pre && !pre->is_null()) {
      *err = "a SentencePiece-style normalizer together with a pre_tokenizer is not supported";
      return false;
    }
  } else if (pre && !pre->is_null()) {
    // byte-level family: ByteLevel alone (GPT-2 / OPT) or a Sequence around it (Falcon: Punctuation(Contiguous),
    // ByteLevel, Digits, Split([0-9][0-9][0-9], Isolated)); each stage re-splits the pieces of the previous one
    std::vector<const Json*> items;
    if (pre->get_str("type", "") == "Sequence") {
      if (const Json* ps = pre->find("pretokenizers"))
        for (auto& q : ps->arr) items.push_back(&q);
    } else {
      items.push_back(pre);
    }
    int n_bl = 0;
    for (const Json* q : items) {
      const std::string t = q->get_str("type", "");
      PreStage st;
      if (t == "ByteLevel") {
        st.kind = PreStage::kByteLevel;
        st.a = q->get_num("add_prefix_space", 1) != 0;
        st.b = q->get_num("use_regex", 1) != 0;
        ++n_bl;
      } else if (t == "Punctuation" || t == "Digits") {
        st.kind = t == "Digits" ? PreStage::kDigits : PreStage::kPunct;
        const std::string beh = q->get_str("behavior", "Isolated");
        if (t == "Digits")
          st.a = q->get_num("individual_digits", 0) == 0;  // a = contiguous
        else if (beh == "Contiguous" || beh == "Isolated")
          st.a = beh == "Contiguous";
        else {
          *err = "unsupported Punctuation behavior '" + beh + "'";
          return false;
        }
      } else if (t == "Split" && q->find("pattern") && q->find("pattern")->get_str("Regex", "") == "[0-9][0-9][0-9]" &&
                 q->get_str("behavior", "") == "Isolated" && q->get_num("invert", 0) == 0) {
        st.kind = PreStage::kSplitDigits3;
      } else {
        *err = "unsupported pre_tokenizer '" + t + "' (ByteLevel, Punctuation, Digits, Split([0-9][0-9][0-9], Isolated))";
        return false;
      }
      pre_.push_back(st);
    }
    if (n_bl != 1) {
      *err = "unsupported pre_tokenizer (exactly one ByteLevel stage expected)";
      return false;
    }
    byte_level_ = true;
  } else {
    *err = "tokenizer has neither a supported normalizer nor a ByteLevel pre_tokenizer";
    return false;
  }
  if (byte_level_) {
    // GPT-2 bytes_to_unicode
    int n = 0;
    for (int b = 0; b < 256; ++b) {
      const bool keep = (b >= 33 && b <= 126) || (b >= 161 && b <= 172) || (b >= 174 && b <= 255);
      std::string u;
      utf8_append(u, keep ? (uint32_t)b : (uint32_t)(256 + n++));
      byte_to_unicode_[b] = u;
      unicode_to_byte_[u] = (uint8_t)b;
    }
  }
  // post processor: TemplateProcessing "single" = [SpecialToken..., Sequence A, SpecialToken...]
  if (const Json* pp = j.find("post_processor"))
    if (!pp->is_null() && pp->get_str("type", "") == "TemplateProcessing")
      if (const Json* single = pp->find("single")) {
        bool seen_a = false;
        for (auto& it : single->arr) {
          if (it.find("Sequence")) {
            seen_a = true;
          } else if (const Json* st = it.find("SpecialToken")) {
            auto v = vocab_.find(st->get_str("id", ""));
            if (v != vocab_.end()) (seen_a ? post_special_ : pre_special_).push_back(v->second);
          }
        }
      }
  return true;
}

// BPE over one "word": initial symbols = its characters (byte fallback for characters missing from the vocab)
void Tokenizer::bpe_word(const std::string& word, std::vector<int32_t>* out) const {
  std::vector<Sym> syms;
  size_t i = 0;
  while (i < word.size()) {
    size_t n;
    utf8_decode(word, i, &n);
    const std::string ch = word.substr(i, n);
    auto it = vocab_.find(ch);
    if (it != vocab_.end()) {
      syms.push_back({it->second, 0, 0});
    } else if (byte_fallback_) {
      for (size_t k = 0; k < n; ++k) {
        const int32_t bt = byte_tokens_[(unsigned char)word[i + k]];
        syms.push_back({bt >= 0 ? bt : unk_id_, 0, 0});
      }
    } else if (unk_id_ >= 0) {
      if (syms.empty() || syms.back().id != unk_id_) syms.push_back({unk_id_, 0, 0});  // fuse_unk
    }
    i += n;
  }
  const int ns = (int)syms.size();
  for (int k = 0; k < ns; ++k) {
    syms[k].prev = k - 1;
    syms[k].next = k + 1 < ns ? k + 1 : -1;
  }
  std::priority_queue<Cand, std::vector<Cand>, std::greater<Cand>> pq;
  auto push = [&](int pos) {
    const int nx = syms[pos].next;
    if (nx < 0) return;
    auto m = merges_.find(((uint64_t)(uint32_t)syms[pos].id << 32) | (uint32_t)syms[nx].id);
    if (m != merges_.end()) pq.push({m->second.first, pos, syms[pos].id, syms[nx].id});
  };
  for (int k = 0; k + 1 < ns; ++k) push(k);
  while (!pq.empty()) {
    const Cand c = pq.top();
    pq.pop();
    if (syms[c.pos].id != c.left) continue;  // stale
    const int nx = syms[c.pos].next;
    if (nx < 0 || syms[nx].id != c.right) continue;
    auto m = merges_.find(((uint64_t)(uint32_t)c.left << 32) | (uint32_t)c.right);
    if (m == merges_.end()) continue;
    syms[c.pos].id = m->second.second;
    syms[nx].id = -1;  // dead
    syms[c.pos].next = syms[nx].next;
    if (syms[nx].next >= 0) syms[syms[nx].next].prev = c.pos;
    if (syms[c.pos].prev >= 0) push(syms[c.pos].prev);
    push(c.pos);
  }
  for (int k = 0; k >= 0 && k < ns; k = syms[k].next) out->push_back(syms[k].id);
}

void Tokenizer::encode_segment(const std::string& text, bool, std::vector<int32_t>* out) const {
  if (text.empty()) return;
  if (!byte_level_) {
    // Llama-2 family: Prepend("▁") then Replace(" ", "▁"); the whole segment is one BPE word (no pre-tokenizer)
    std::string w;
    if (sp_prepend_) w += "\xE2\x96\x81";
    for (char c : text) {
      if (c == ' ')
        w += "\xE2\x96\x81";
      else
        w += c;
    }
    bpe_word(w, out);
    return;
  }
  // byte-level family: run the pre-tokenizer stages over the growing list of pieces, BPE every final piece
  std::vector<std::string> pieces{text}, next;
  for (const PreStage& st : pre_) {
    next.clear();
    for (const std::string& pc : pieces) {
      switch (st.kind) {
        case PreStage::kByteLevel: {
          std::string t = pc;
          if (st.a && t[0] != ' ') t = " " + t;
          std::vector<std::string> raw;
          if (st.b)
            gpt2_split(t, &raw);
          else
            raw.push_back(t);
          for (const std::string& r : raw) {
            std::string mapped;
            for (unsigned char ch : r) mapped += byte_to_unicode_[ch];
            next.push_back(mapped);
          }
          break;
        }
        case PreStage::kPunct: split_chars(pc, is_punct, st.a, &next); break;
        case PreStage::kDigits: split_chars(pc, is_numeric, st.a, &next); break;
        case PreStage::kSplitDigits3: {
          auto dig = [&](size_t k) { return k < pc.size() && pc[k] >= '0' && pc[k] <= '9'; };
          size_t start = 0, i = 0;
          while (i < pc.size()) {
            if (dig(i) && dig(i + 1) && dig(i + 2)) {
              if (i > start) next.push_back(pc.substr(start, i - start));
              next.push_back(pc.substr(i, 3));
              i += 3;
              start = i;
            } else {
              ++i;
            }
          }
          if (start < pc.size()) next.push_back(pc.substr(start));
          break;
        }
      }
    }
    pieces.swap(next);
  }
  for (const std::string& pc : pieces) bpe_word(pc, out);
}

// The GPT-2 pattern over one piece of text (raw bytes out; lookaheads see the end of the piece as end of input).
void Tokenizer::gpt2_split(const std::string& t, std::vector<std::string>* pieces) {
  std::vector<uint32_t> cps;
  std::vector<size_t> offs;
  for (size_t i = 0; i < t.size();) {
    size_t n;
    cps.push_back(utf8_decode(t, i, &n));
    offs.push_back(i);
    i += n;
  }
  offs.push_back(t.size());
  const size_t N = cps.size();
  auto emit = [&](size_t a, size_t b) { pieces->push_back(t.substr(offs[a], offs[b] - offs[a])); };
  size_t p = 0;
  while (p < N) {
    // 1. contractions
    if (cps[p] == '\'' && p + 1 < N) {
      static const char* two[] = {"re", "ve", "ll"};
      const uint32_t c1 = cps[p + 1];
      bool m2 = false;
      if (p + 2 < N)
        for (auto s : two)
          if (c1 == (uint32_t)s[0] && cps[p + 2] == (uint32_t)s[1]) m2 = true;
      if (c1 == 's' || c1 == 't' || c1 == 'm' || c1 == 'd') {
        emit(p, p + 2);
        p += 2;
        continue;
      }
      if (m2) {
        emit(p, p + 3);
        p += 3;
        continue;
      }
    }
    // 2-4. optional single space + run of letters / numbers / other non-space
    size_t q = p;
    if (cps[q] == ' ' && q + 1 < N) ++q;
    auto cls = [&](uint32_t c) { return is_letter(c) ? 1 : is_number(c) ? 2 : is_space(c) ? 0 : 3; };
    const int k = q < N ? cls(cps[q]) : 0;
    if (k != 0) {
      size_t e = q;
      while (e < N && cls(cps[e]) == k) ++e;
      emit(p, e);
      p = e;
      continue;
    }
    // 5-6. whitespace run: all but the last char if a non-space follows (the last one joins the next piece when it is
    // a plain ' '; otherwise it is emitted alone by the next iteration through this same branch)
    size_t e = p;
    while (e < N && is_space(cps[e])) ++e;
    if (e < N && e - p > 1) --e;
    emit(p, e);
    p = e;
  }
}

std::vector<int32_t> Tokenizer::encode(const std::string& text, bool add_special) const {
  std::vector<int32_t> out;
  if (add_special) out = pre_special_;
  // split on added tokens (matched verbatim, longest first, leftmost)
  size_t pos = 0, seg_start = 0;
  bool first = true;
  while (pos < text.size() && !added_.empty()) {
    int hit = -1;
    for (size_t a = 0; a < added_.size(); ++a)
      if (text.compare(pos, added_[a].first.size(), added_[a].first) == 0) {
        hit = (int)a;
        break;
      }
    if (hit < 0) {
      ++pos;
      continue;
    }
    encode_segment(text.substr(seg_start, pos - seg_start), first, &out);
    first = false;
    out.push_back(added_[hit].second);
    pos += added_[hit].first.size();
    seg_start = pos;
  }
  encode_segment(text.substr(seg_start), first, &out);
  if (add_special) out.insert(out.end(), post_special_.begin(), post_special_.end());
  return out;
}

std::string Tokenizer::decode(const std::vector<int32_t>& ids, bool skip_special) const {
  std::string out;
  if (byte_level_) {
    std::string bytes;
    for (int32_t id : ids) {
      if (id < 0 || (size_t)id >= id_to_token_.size()) continue;
      if (skip_special && is_special_[id]) continue;
      const std::string& tk = id_to_token_[id];
      if (is_special_[id]) {
        bytes += tk;
        continue;
      }
      for (size_t i = 0; i < tk.size();) {
        size_t n;
        utf8_decode(tk, i, &n);
        auto it = unicode_to_byte_.find(tk.substr(i, n));
        if (it != unicode_to_byte_.end())
          bytes += (char)it->second;
        else
          bytes += tk.substr(i, n);
        i += n;
      }
    }
    if (utf8_valid(bytes)) return bytes;
    // lossy: keep well-formed sequences, one U+FFFD per maximal ill-formed prefix (String::from_utf8_lossy)
    for (size_t i = 0; i < bytes.size();) {
      const unsigned char c = (unsigned char)bytes[i];
      const size_t len = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 0;
      if (len && i + len <= bytes.size() && utf8_valid(bytes.substr(i, len))) {
        out += bytes.substr(i, len);
        i += len;
      } else {
        out += "\xEF\xBF\xBD";
        size_t k = 1;
        while (i + k < bytes.size() && ((unsigned char)bytes[i + k] & 0xC0) == 0x80 && k < (len ? len : 1)) ++k;
        i += k;
      }
    }
    return out;
  }
  // Llama-2 family: Replace("▁", " "), ByteFallback, Fuse, Strip(1 leading space)
  std::string pending;  // consecutive <0xXX> tokens
  size_t n_pending = 0;
  auto flush = [&]() {
    if (!n_pending) return;
    if (utf8_valid(pending))
      out += pending;
    else
      for (size_t k = 0; k < n_pending; ++k) out += "\xEF\xBF\xBD";
    pending.clear();
    n_pending = 0;
  };
  for (int32_t id : ids) {
    if (id < 0 || (size_t)id >= id_to_token_.size()) continue;
    if (skip_special && is_special_[id]) continue;
    const std::string& tk = id_to_token_[id];
    if (!is_special_[id] && tk.size() == 6 && tk.compare(0, 3, "<0x") == 0 && tk[5] == '>') {
      pending += (char)strtol(tk.substr(3, 2).c_str(), nullptr, 16);
      ++n_pending;
      continue;
    }
    flush();
    for (size_t i = 0; i < tk.size();) {
      if (tk.compare(i, 3, "\xE2\x96\x81") == 0) {
        out += ' ';
        i += 3;
      } else {
        out += tk[i++];
      }
    }
  }
  flush();
  if (!out.empty() && out[0] == ' ') out.erase(0, 1);
  return out;
}

}  // namespace ssb
