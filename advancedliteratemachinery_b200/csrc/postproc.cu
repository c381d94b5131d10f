// Host-only post-processing behind the C ABI: decoded sequences -> structured results / JSON (OmniParser) and the
// three-head confidence fusion (MGP-STR).  No CUDA calls, no context: these entry points work on a machine without
// a GPU and are parity-tested on the CPU against outputs of the reference's own functions (tests/golden/post_*.json).
//
// Reference being replaced: OCR/OmniParser/engine/val.py:63-100, utils/misc.py:147-189;
//                           OCR/MGP-STR/test_final.py:176-240, demo.py:36-112, utils.py:52-87.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <charconv>
#include <string>
#include <vector>

#include "../../include/alm_ocr.h"

namespace {

thread_local std::string g_post_err;

struct PostError {
  int code;
  std::string msg;
};

#define POST_REQUIRE(cond, message)                                        \
  do {                                                                     \
    if (!(cond)) throw PostError{ALM_ERR_INVALID, std::string(message)};   \
  } while (0)

template <class F>
int post_guard(F&& f) {
  try {
    f();
    return ALM_OK;
  } catch (const PostError& e) {
    g_post_err = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    g_post_err = e.what();
    return ALM_ERR_INVALID;
  }
}

// ---------------------------------------------------------------------------------------------- UTF-8
// code points of a UTF-8 string as byte ranges (malformed bytes count as one code point each)
std::vector<std::pair<size_t, size_t>> utf8_spans(const std::string& s) {
  std::vector<std::pair<size_t, size_t>> out;
  for (size_t i = 0; i < s.size();) {
    const unsigned char c = static_cast<unsigned char>(s[i]);
    size_t n = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 1;
    if (i + n > s.size()) n = 1;
    out.emplace_back(i, n);
    i += n;
  }
  return out;
}

uint32_t utf8_decode(const char* p, size_t n) {
  const unsigned char* u = reinterpret_cast<const unsigned char*>(p);
  if (n == 1) return u[0];
  if (n == 2) return ((u[0] & 0x1Fu) << 6) | (u[1] & 0x3Fu);
  if (n == 3) return ((u[0] & 0x0Fu) << 12) | ((u[1] & 0x3Fu) << 6) | (u[2] & 0x3Fu);
  return ((u[0] & 0x07u) << 18) | ((u[1] & 0x3Fu) << 12) | ((u[2] & 0x3Fu) << 6) | (u[3] & 0x3Fu);
}

// Python str.find on code points: index of the first occurrence of the (ASCII) needle, or -1
long py_find(const std::string& s, const std::vector<std::pair<size_t, size_t>>& spans, const char* needle) {
  const size_t pos = s.find(needle);
  if (pos == std::string::npos) return -1;
  long idx = 0;
  for (const auto& sp : spans) {
    if (sp.first >= pos) break;
    ++idx;
  }
  return idx;
}

// Python s[:k] on code points (negative k counts from the end, like Python)
std::string py_prefix(const std::string& s, const std::vector<std::pair<size_t, size_t>>& spans, long k) {
  const long n = static_cast<long>(spans.size());
  if (k < 0) k = std::max<long>(0, n + k);
  if (k >= n) return s;
  return s.substr(0, spans[static_cast<size_t>(k)].first);
}

// ---------------------------------------------------------------------------------------------- numbers
// repr(float) of CPython (float_repr_style 'short'): shortest digits that round-trip, fixed notation for
// 1e-4 <= |x| < 1e16, exponent notation otherwise, always a '.0' on integral fixed values.
std::string py_float_repr(double v) {
  if (std::isnan(v)) return "NaN";          // json.dumps spelling (allow_nan=True)
  if (std::isinf(v)) return v > 0 ? "Infinity" : "-Infinity";
  if (v == 0.0) return std::signbit(v) ? "-0.0" : "0.0";
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);  // shortest round-trip digits
  std::string sci(buf, r.ptr);
  std::string out;
  size_t i = 0;
  if (sci[0] == '-') { out = "-"; i = 1; }
  const size_t epos = sci.find('e');
  std::string digits;
  for (size_t k = i; k < epos; ++k)
    if (sci[k] != '.') digits.push_back(sci[k]);
  const int exp10 = std::stoi(sci.substr(epos + 1));
  const int nd = static_cast<int>(digits.size());
  if (exp10 < -4 || exp10 >= 16) {
    out += digits[0];
    if (nd > 1) { out += '.'; out += digits.substr(1); }
    out += 'e';
    out += exp10 < 0 ? '-' : '+';
    const int a = exp10 < 0 ? -exp10 : exp10;
    if (a < 10) out += '0';
    out += std::to_string(a);
    return out;
  }
  if (exp10 < 0) {
    out += "0.";
    out.append(static_cast<size_t>(-exp10 - 1), '0');
    out += digits;
  } else if (nd <= exp10 + 1) {
    out += digits;
    out.append(static_cast<size_t>(exp10 + 1 - nd), '0');
    out += ".0";
  } else {
    out += digits.substr(0, static_cast<size_t>(exp10 + 1));
    out += '.';
    out += digits.substr(static_cast<size_t>(exp10 + 1));
  }
  return out;
}

// json.dumps string literal with ensure_ascii=True
std::string json_string(const std::string& s) {
  static const char* hex = "0123456789abcdef";
  std::string out = "\"";
  auto u_escape = [&](uint32_t cp) {
    out += "\\u";
    out += hex[(cp >> 12) & 15]; out += hex[(cp >> 8) & 15]; out += hex[(cp >> 4) & 15]; out += hex[cp & 15];
  };
  for (const auto& sp : utf8_spans(s)) {
    const uint32_t cp = utf8_decode(s.data() + sp.first, sp.second);
    if (sp.second == 1 && cp < 0x80) {
      const char c = static_cast<char>(cp);
      switch (c) {
        case '"': out += "\\\""; break;
        case '\\': out += "\\\\"; break;
        case '\n': out += "\\n"; break;
        case '\r': out += "\\r"; break;
        case '\t': out += "\\t"; break;
        case '\b': out += "\\b"; break;
        case '\f': out += "\\f"; break;
        default:
          if (cp < 0x20) u_escape(cp);
          else out += c;
      }
    } else if (cp >= 0x10000) {
      const uint32_t v = cp - 0x10000;
      u_escape(0xD800 + (v >> 10));
      u_escape(0xDC00 + (v & 0x3FF));
    } else {
      u_escape(cp);
    }
  }
  out += '"';
  return out;
}

// ---------------------------------------------------------------------------------------------- OmniParser
struct SpotInstance {
  double pt[2];
  double poly[32];
  double score;
  std::string rec;
};

// torch: (int64 tensor / python int) -> float32 true division; (* int64 size) -> float32 product
inline float bin_to_coord(int64_t id, int num_bins, long size) {
  const float q = static_cast<float>(id) / static_cast<float>(num_bins);
  return q * static_cast<float>(size);
}

std::vector<SpotInstance> spot_instances(const int64_t* pt, const int64_t* poly, const int64_t* rec, const float* rec_prob,
                                         int n, int rec_length, int num_bins, int recog_pad_index, int rec_eos_index,
                                         const char* chars, long orig_h, long orig_w) {
  POST_REQUIRE(n >= 0 && rec_length > 0 && num_bins > 0 && chars, "alm_post_omni: arguments");
  POST_REQUIRE(n == 0 || (pt && poly && rec && rec_prob), "alm_post_omni: null sequence buffer");
  const std::string cs(chars);
  const auto spans = utf8_spans(cs);
  std::vector<SpotInstance> out(static_cast<size_t>(n));
  for (int i = 0; i < n; ++i) {
    SpotInstance& r = out[static_cast<size_t>(i)];
    // decode_seq 'pt' (misc.py:149-155) then val.py:88-89
    r.pt[0] = bin_to_coord(pt[2 * i], num_bins, orig_w);
    r.pt[1] = bin_to_coord(pt[2 * i + 1], num_bins, orig_h);
    // decode_seq 'poly' (misc.py:156-161) then val.py:90: * tensor([w, h] * 16)
    for (int k = 0; k < 32; ++k) r.poly[k] = bin_to_coord(poly[32 * i + k], num_bins, (k & 1) ? orig_h : orig_w);
    // decode_seq 'rec' (misc.py:162-184): stop at pad / eos, skip 'unknown', mean kept probability
    double psum = 0.0;
    int kept = 0;
    for (int j = 0; j < rec_length; ++j) {
      const int64_t id = rec[static_cast<size_t>(i) * rec_length + j];
      if (id == recog_pad_index || id == rec_eos_index) break;
      if (id == recog_pad_index - 1) continue;
      int64_t ci = id - num_bins;
      const int64_t ncs = static_cast<int64_t>(spans.size());
      if (ci < 0) ci += ncs;  // Python negative index
      POST_REQUIRE(ci >= 0 && ci < ncs, "alm_post_omni: recognition id " + std::to_string(id) + " has no character (IndexError in the reference)");
      r.rec.append(cs, spans[static_cast<size_t>(ci)].first, spans[static_cast<size_t>(ci)].second);
      psum += static_cast<double>(rec_prob[static_cast<size_t>(i) * rec_length + j]);
      ++kept;
    }
    r.score = psum / (static_cast<double>(kept) + 1e-5);
  }
  return out;
}

void indent(std::string& s, int level) { s.append(static_cast<size_t>(4 * level), ' '); }

std::string spot_json(const std::vector<SpotInstance>& inst, const std::string& image_id) {
  if (inst.empty()) return "[]";
  std::string s = "[\n";
  for (size_t i = 0; i < inst.size(); ++i) {
    const SpotInstance& r = inst[i];
    indent(s, 1); s += "{\n";
    indent(s, 2); s += "\"image_id\": " + json_string(image_id) + ",\n";
    indent(s, 2); s += "\"pts\": [\n";
    indent(s, 3); s += "[\n";
    indent(s, 4); s += py_float_repr(r.pt[0]) + ",\n";
    indent(s, 4); s += py_float_repr(r.pt[1]) + "\n";
    indent(s, 3); s += "]\n";
    indent(s, 2); s += "],\n";
    indent(s, 2); s += "\"score\": " + py_float_repr(r.score) + ",\n";
    indent(s, 2); s += "\"polys\": [\n";
    for (int k = 0; k < 16; ++k) {
      indent(s, 3); s += "[\n";
      indent(s, 4); s += py_float_repr(r.poly[2 * k]) + ",\n";
      indent(s, 4); s += py_float_repr(r.poly[2 * k + 1]) + "\n";
      indent(s, 3); s += k == 15 ? "]\n" : "],\n";
    }
    indent(s, 2); s += "],\n";
    indent(s, 2); s += "\"rec\": " + json_string(r.rec) + "\n";
    indent(s, 1); s += i + 1 == inst.size() ? "}\n" : "},\n";
  }
  s += "]";
  return s;
}

// ---------------------------------------------------------------------------------------------- MGP-STR
// PreTrainedTokenizerBase.clean_up_tokenization (transformers 4.2.1, the reference's pin: decode() applies it by
// default): the same global replacements in the same order
void replace_all(std::string& s, const char* from, const char* to) {
  const size_t nf = strlen(from), nt = strlen(to);
  for (size_t pos = s.find(from); pos != std::string::npos; pos = s.find(from, pos + nt)) s.replace(pos, nf, to);
}
void clean_up_tokenization(std::string& s) {
  replace_all(s, " .", ".");   replace_all(s, " ?", "?");     replace_all(s, " !", "!");   replace_all(s, " ,", ",");
  replace_all(s, " ' ", "'");  replace_all(s, " n't", "n't"); replace_all(s, " 'm", "'m"); replace_all(s, " 's", "'s");
  replace_all(s, " 've", "'ve"); replace_all(s, " 're", "'re");
}

struct HeadResult {
  std::string text;
  float conf;
};

// confidence = float32(prod in double of p[0..k)) -- torch.cumprod accumulates float32 tensors in double on the CPU
float cumprod_last(const float* p, long k) {
  if (k <= 0) return 0.0f;  // empty slice -> cumprod()[-1] raises -> the reference's `except: 0.0`
  double acc = 1.0;
  for (long i = 0; i < k; ++i) acc *= static_cast<double>(p[i]);
  return static_cast<float>(acc);
}

}  // namespace

extern "C" {

const char* alm_post_last_error(void) { return g_post_err.c_str(); }

int alm_post_omni_spotting(const int64_t* pt, const int64_t* poly, const int64_t* rec, const float* rec_prob, int n,
                           int rec_length, int num_bins, int recog_pad_index, int rec_eos_index, const char* chars,
                           long orig_h, long orig_w, double* pts, double* polys, double* scores, char* texts,
                           size_t text_stride) {
  return post_guard([&] {
    const auto inst = spot_instances(pt, poly, rec, rec_prob, n, rec_length, num_bins, recog_pad_index, rec_eos_index, chars,
                                     orig_h, orig_w);
    POST_REQUIRE(n == 0 || (pts && polys && scores && texts), "alm_post_omni_spotting: null output buffer");
    for (int i = 0; i < n; ++i) {
      const SpotInstance& r = inst[static_cast<size_t>(i)];
      pts[2 * i] = r.pt[0]; pts[2 * i + 1] = r.pt[1];
      for (int k = 0; k < 32; ++k) polys[32 * i + k] = r.poly[k];
      scores[i] = r.score;
      POST_REQUIRE(r.rec.size() + 1 <= text_stride, "alm_post_omni_spotting: text_stride too small");
      memcpy(texts + static_cast<size_t>(i) * text_stride, r.rec.c_str(), r.rec.size() + 1);
    }
  });
}

int alm_post_omni_json(const int64_t* pt, const int64_t* poly, const int64_t* rec, const float* rec_prob, int n,
                       int rec_length, int num_bins, int recog_pad_index, int rec_eos_index, const char* chars, long orig_h,
                       long orig_w, const char* image_id, char* json, size_t cap, size_t* needed) {
  return post_guard([&] {
    POST_REQUIRE(image_id && needed, "alm_post_omni_json: arguments");
    const std::string s = spot_json(spot_instances(pt, poly, rec, rec_prob, n, rec_length, num_bins, recog_pad_index,
                                                   rec_eos_index, chars, orig_h, orig_w),
                                    image_id);
    *needed = s.size() + 1;
    POST_REQUIRE(json && cap >= s.size() + 1, "alm_post_omni_json: buffer too small (see *needed)");
    memcpy(json, s.c_str(), s.size() + 1);
  });
}

int alm_post_omni_kie_json(const int64_t* tokens, const float* probs, int n_tok, const int32_t* inst_pos, int n_inst,
                           const int64_t* poly, const int64_t* rec, int rec_length, int num_bins, int recog_pad_index,
                           int rec_eos_index, const char* chars, const char* const* classes, int n_classes, int class_base,
                           long orig_h, long orig_w, char* json, size_t cap, size_t* needed) {
  return post_guard([&] {
    POST_REQUIRE(n_tok >= 0 && n_inst >= 0 && rec_length > 0 && num_bins > 0 && chars && classes && needed,
                 "alm_post_omni_kie_json: arguments");
    POST_REQUIRE(n_tok == 0 || (tokens && probs), "alm_post_omni_kie_json: null token buffers");
    POST_REQUIRE(n_inst == 0 || (inst_pos && poly && rec), "alm_post_omni_kie_json: null instance buffers");
    const std::string cs(chars);
    const auto spans = utf8_spans(cs);
    // the walk of decode_vie_pt_poly_rec_seq (transformer.py:148-215): an (x, y) pair of bins is a word (its polygon
    // extent and transcription were decoded by the batched poly / rec loops), a lone bin is skipped, any other token
    // is an entity class and closes the entity collected so far
    std::string out = "[";
    std::vector<std::string> words;
    std::vector<double> rects;
    bool first_entity = true;
    int next_inst = 0;
    for (int i = 0; i < n_tok;) {
      const int64_t tk = tokens[i];
      if (tk < num_bins) {
        if (i + 1 <= n_tok - 1 && tokens[i + 1] < num_bins) {
          POST_REQUIRE(next_inst < n_inst && inst_pos[next_inst] == i,
                       "alm_post_omni_kie_json: instance list does not match the (x, y) pairs of the token stream");
          const int64_t* pp = poly + static_cast<size_t>(next_inst) * 32;
          int64_t mnx = pp[0], mxx = pp[0], mny = pp[1], mxy = pp[1];
          for (int k = 0; k < 16; ++k) {
            mnx = std::min(mnx, pp[2 * k]); mxx = std::max(mxx, pp[2 * k]);
            mny = std::min(mny, pp[2 * k + 1]); mxy = std::max(mxy, pp[2 * k + 1]);
          }
          // image_w.item() * min.item() / num_bins: exact integer product, then one double division (:165-168)
          rects.push_back(static_cast<double>(orig_w * mnx) / num_bins);
          rects.push_back(static_cast<double>(orig_h * mny) / num_bins);
          rects.push_back(static_cast<double>(orig_w * mxx) / num_bins);
          rects.push_back(static_cast<double>(orig_h * mxy) / num_bins);
          std::string w;
          for (int j = 0; j < rec_length; ++j) {
            const int64_t id = rec[static_cast<size_t>(next_inst) * rec_length + j];
            if (id == recog_pad_index || id == rec_eos_index) break;
            if (id == recog_pad_index - 1) continue;
            int64_t ci = id - num_bins;
            const int64_t ncs = static_cast<int64_t>(spans.size());
            if (ci < 0) ci += ncs;
            POST_REQUIRE(ci >= 0 && ci < ncs, "alm_post_omni_kie_json: recognition id has no character");
            w.append(cs, spans[static_cast<size_t>(ci)].first, spans[static_cast<size_t>(ci)].second);
          }
          words.push_back(w);
          ++next_inst;
          i += 2;
        } else {
          i += 1;
        }
        continue;
      }
      const int64_t cls = tk - class_base;
      POST_REQUIRE(cls >= 0 && cls < n_classes, "alm_post_omni_kie_json: token is neither a bin nor an entity class (KeyError in the reference)");
      std::string text;
      for (size_t k = 0; k < words.size(); ++k) {
        if (k) text += ' ';
        text += words[k];
      }
      if (!first_entity) out += ", ";
      first_entity = false;
      out += "[" + json_string(text) + ", " + json_string(classes[cls]) + ", " + py_float_repr(static_cast<double>(probs[i])) + ", [";
      for (size_t k = 0; k < rects.size(); k += 4) {
        if (k) out += ", ";
        out += "[" + py_float_repr(rects[k]) + ", " + py_float_repr(rects[k + 1]) + ", " + py_float_repr(rects[k + 2]) + ", " +
               py_float_repr(rects[k + 3]) + "]";
      }
      out += "]]";
      words.clear();
      rects.clear();
      i += 1;
    }
    out += "]";
    *needed = out.size() + 1;
    POST_REQUIRE(json && cap >= out.size() + 1, "alm_post_omni_kie_json: buffer too small (see *needed)");
    memcpy(json, out.c_str(), out.size() + 1);
  });
}

int alm_post_mgp_fuse(const int32_t* ids, const float* prob, int B, int T, const char* const* char_table, int n_char,
                      const char* const* bpe_table, int n_bpe, const char* const* wp_table, int n_wp, char* texts,
                      char* fused, size_t text_stride, float* conf, int32_t* source) {
  return post_guard([&] {
    POST_REQUIRE(ids && prob && B >= 0 && T >= 2 && char_table && bpe_table && wp_table && texts && fused && conf && source,
                 "alm_post_mgp_fuse: arguments");
    const char* const* tables[3] = {char_table, bpe_table, wp_table};
    const int sizes[3] = {n_char, n_bpe, n_wp};
    const char* eos_str[3] = {"[s]", "#", "[SEP]"};
    const int eos_id[3] = {1, 2, 102};
    auto put = [&](char* dst, const std::string& s) {
      POST_REQUIRE(s.size() + 1 <= text_stride, "alm_post_mgp_fuse: text_stride too small");
      memcpy(dst, s.c_str(), s.size() + 1);
    };
    for (int b = 0; b < B; ++b) {
      HeadResult res[3];
      for (int hd = 0; hd < 3; ++hd) {
        const int32_t* id = ids + (static_cast<size_t>(hd) * B + b) * T + 1;  // positions 1..T-1 ([:, 1:])
        const float* p = prob + (static_cast<size_t>(hd) * B + b) * T + 1;
        const int L = T - 1;
        // token strings -> text (utils.py:52-87)
        std::string s;
        for (int j = 0; j < L; ++j) {
          POST_REQUIRE(id[j] >= 0 && id[j] < sizes[hd], "alm_post_mgp_fuse: token id outside its table");
          const char* tok = tables[hd][id[j]];
          if (hd == 2 && j > 0 && tok[0] == '#' && tok[1] == '#') tok += 2;  // BERT decode joins '##' pieces; the
          s += tok;                                                         // reference then drops all whitespace
        }
        if (hd == 1) clean_up_tokenization(s);  // GPT2Tokenizer.decode (utils.py:73)
        if (hd == 2) {  // ''.join(tokenstr.split()): whitespace inside tokens goes too
          std::string t;
          for (char ch : s)
            if (!(ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r' || ch == '\f' || ch == '\v')) t += ch;
          s.swap(t);
        }
        const auto spans = utf8_spans(s);
        const long eos = py_find(s, spans, eos_str[hd]);
        res[hd].text = py_prefix(s, spans, eos);  // eos == -1 drops the last character, like the reference
        long k;
        if (hd == 0) {
          k = eos + 1;  // the char head slices the probabilities with the STRING index (test_final.py:188)
        } else {
          long idx = -1;
          for (int j = 0; j < L; ++j)
            if (id[j] == eos_id[hd]) { idx = j; break; }
          k = idx + 1;  // test_final.py:203-207 / :222-226
        }
        res[hd].conf = cumprod_last(p, std::min<long>(k, L));
        put(texts + (static_cast<size_t>(hd) * B + b) * text_stride, res[hd].text);
        conf[static_cast<size_t>(hd) * B + b] = res[hd].conf;
      }
      float best = 0.0f;
      int src = -1;
      for (int hd = 0; hd < 3; ++hd)
        if (res[hd].conf > best) { best = res[hd].conf; src = hd; }
      source[b] = src;
      put(fused + static_cast<size_t>(b) * text_stride, src < 0 ? std::string() : res[src].text);
    }
  });
}

}  // extern "C"
