// Protobuf text-format reader, just enough for config/bayesian_segnet/*/kitti/*.prototxt.
#include "prototxt.h"

#include <cctype>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>

#include "common.h"

namespace sivo {
namespace {

struct Msg;
struct Value {
  bool is_msg = false, blank = false;
  std::string s;
  std::shared_ptr<Msg> m;
};
struct Msg {
  std::vector<std::pair<std::string, Value>> fields;
  std::vector<const Value*> all(const std::string& k) const {
    std::vector<const Value*> r;
    for (auto& f : fields)
      if (f.first == k) r.push_back(&f.second);
    return r;
  }
  const Value* one(const std::string& k) const {
    for (auto& f : fields)
      if (f.first == k) return &f.second;
    return nullptr;
  }
  std::string str(const std::string& k, const std::string& d = "") const {
    auto* v = one(k);
    return (v && !v->is_msg && !v->blank) ? v->s : d;
  }
  const Msg* msg(const std::string& k) const {
    auto* v = one(k);
    return (v && v->is_msg) ? v->m.get() : nullptr;
  }
};

struct Tok {
  enum Kind { Word, Str, Punct, LineEnd, End } kind;
  std::string s;
};

class Lexer {
 public:
  explicit Lexer(const std::string& t) : t_(t) {}
  Tok next() {
    while (p_ < t_.size()) {
      char c = t_[p_];
      if (c == '#') {  // comment: ends a scalar, so `dim: # ...` is seen as blank
        while (p_ < t_.size() && t_[p_] != '\n') ++p_;
        return {Tok::LineEnd, ""};
      }
      if (isspace(static_cast<unsigned char>(c))) { ++p_; continue; }
      if (c == '{' || c == '}' || c == ':') { ++p_; return {Tok::Punct, std::string(1, c)}; }
      if (c == '"' || c == '\'') {
        char q = c;
        size_t b = ++p_;
        while (p_ < t_.size() && t_[p_] != q) p_ += (t_[p_] == '\\') ? 2 : 1;
        if (p_ >= t_.size()) fail(SIVO_EFORMAT, "prototxt: unterminated string");
        std::string s = t_.substr(b, p_ - b);
        ++p_;
        return {Tok::Str, s};
      }
      size_t b = p_;
      while (p_ < t_.size() && !isspace(static_cast<unsigned char>(t_[p_])) && t_[p_] != '{' && t_[p_] != '}' &&
             t_[p_] != ':' && t_[p_] != '#' && t_[p_] != '"')
        ++p_;
      return {Tok::Word, t_.substr(b, p_ - b)};
    }
    return {Tok::End, ""};
  }

 private:
  const std::string& t_;
  size_t p_ = 0;
};

class Parser {
 public:
  explicit Parser(const std::string& t) : lex_(t) { advance(); }
  std::shared_ptr<Msg> message(int depth) {
    auto m = std::make_shared<Msg>();
    for (;;) {
      while (cur_.kind == Tok::LineEnd) advance();
      if (cur_.kind == Tok::End) {
        if (depth) fail(SIVO_EFORMAT, "prototxt: missing '}'");
        return m;
      }
      if (cur_.kind == Tok::Punct && cur_.s == "}") {
        if (!depth) fail(SIVO_EFORMAT, "prototxt: unbalanced '}'");
        advance();
        return m;
      }
      if (cur_.kind != Tok::Word) fail(SIVO_EFORMAT, "prototxt: expected a field name near '%s'", cur_.s.c_str());
      std::string name = cur_.s;
      advance();
      Value v;
      if (cur_.kind == Tok::Punct && cur_.s == ":") {
        advance();
        if (cur_.kind == Tok::Punct && cur_.s == "{") {
          advance();
          v.is_msg = true;
          v.m = message(depth + 1);
        } else if (cur_.kind == Tok::Word || cur_.kind == Tok::Str) {
          v.s = cur_.s;
          advance();
        } else {
          v.blank = true;  // `dim: # SET SAMPLE SIZE HERE`
        }
      } else if (cur_.kind == Tok::Punct && cur_.s == "{") {
        advance();
        v.is_msg = true;
        v.m = message(depth + 1);
      } else {
        fail(SIVO_EFORMAT, "prototxt: expected ':' or '{' after '%s'", name.c_str());
      }
      m->fields.emplace_back(name, std::move(v));
    }
  }

 private:
  void advance() { cur_ = lex_.next(); }
  Lexer lex_;
  Tok cur_;
};

int to_int(const std::string& s, const char* what) {
  try {
    size_t pos = 0;
    int v = std::stoi(s, &pos);
    if (pos != s.size()) throw std::invalid_argument(s);
    return v;
  } catch (...) {
    fail(SIVO_EFORMAT, "prototxt: '%s' is not an integer (%s)", s.c_str(), what);
  }
}
float to_float(const std::string& s, const char* what) {
  try {
    return std::stof(s);
  } catch (...) {
    fail(SIVO_EFORMAT, "prototxt: '%s' is not a number (%s)", s.c_str(), what);
  }
}
bool to_bool(const std::string& s) { return s == "true" || s == "1" || s == "True"; }

}  // namespace

NetSpec parse_prototxt_text(const std::string& text) {
  Parser p(text);
  auto root = p.message(0);
  NetSpec net;
  net.name = root->str("name");
  net.input_name = root->str("input", "data");
  std::vector<const Value*> dims = root->all("input_dim");
  if (dims.empty())
    if (auto* shp = root->msg("input_shape")) dims = shp->all("dim");
  if (dims.size() == 3) {  // blank first dim swallowed
    net.dims[0] = 0;
    for (int i = 0; i < 3; ++i) net.dims[i + 1] = to_int(dims[i]->s, "input dim");
  } else if (dims.size() == 4) {
    for (int i = 0; i < 4; ++i) net.dims[i] = dims[i]->blank ? 0 : to_int(dims[i]->s, "input dim");
  } else {
    fail(SIVO_EFORMAT, "prototxt: expected 4 input dims, found %zu", dims.size());
  }
  if (!root->all("layers").empty()) fail(SIVO_EFORMAT, "prototxt: V1 'layers' are not supported");
  for (auto* lv : root->all("layer")) {
    if (!lv->is_msg) fail(SIVO_EFORMAT, "prototxt: 'layer' must be a message");
    const Msg& lm = *lv->m;
    LayerSpec ly;
    ly.name = lm.str("name");
    std::string type = lm.str("type");
    for (auto* b : lm.all("bottom")) ly.bottoms.push_back(b->s);
    for (auto* t : lm.all("top")) ly.tops.push_back(t->s);
    if (ly.bottoms.empty() || ly.tops.empty()) fail(SIVO_EFORMAT, "layer '%s' needs a bottom and a top", ly.name.c_str());
    if (type == "Convolution") {
      ly.type = LayerType::Convolution;
      const Msg* cp = lm.msg("convolution_param");
      if (!cp) fail(SIVO_EFORMAT, "layer '%s': missing convolution_param", ly.name.c_str());
      ly.num_output = to_int(cp->str("num_output", "0"), "num_output");
      ly.kernel = to_int(cp->str("kernel_size", "0"), "kernel_size");
      ly.pad = to_int(cp->str("pad", "0"), "pad");
      ly.bias_term = to_bool(cp->str("bias_term", "true"));
      if (to_int(cp->str("stride", "1"), "stride") != 1 || to_int(cp->str("group", "1"), "group") != 1 ||
          to_int(cp->str("dilation", "1"), "dilation") != 1)
        fail(SIVO_EFORMAT, "layer '%s': only stride-1, ungrouped, undilated convolutions are on the path", ly.name.c_str());
      if (ly.num_output <= 0 || ly.kernel <= 0 || (ly.kernel & 1) == 0 || ly.pad != (ly.kernel - 1) / 2)
        fail(SIVO_EFORMAT, "layer '%s': expected an odd kernel with 'same' padding", ly.name.c_str());
    } else if (type == "ReLU") {
      ly.type = LayerType::ReLU;
      if (auto* rp = lm.msg("relu_param")) ly.negative_slope = to_float(rp->str("negative_slope", "0"), "negative_slope");
    } else if (type == "BN") {
      ly.type = LayerType::BN;
      const Msg* bp = lm.msg("bn_param");
      if (!bp || bp->str("bn_mode", "LEARN") != "INFERENCE")
        fail(SIVO_EFORMAT, "layer '%s': BN must be bn_mode INFERENCE", ly.name.c_str());
    } else if (type == "LRN") {
      ly.type = LayerType::LRN;
      if (auto* lp = lm.msg("lrn_param")) {
        ly.local_size = to_int(lp->str("local_size", "5"), "local_size");
        ly.alpha = to_float(lp->str("alpha", "1"), "alpha");
        ly.beta = to_float(lp->str("beta", "0.75"), "beta");
        ly.k = to_float(lp->str("k", "1"), "k");
        if (lp->str("norm_region", "ACROSS_CHANNELS") != "ACROSS_CHANNELS")
          fail(SIVO_EFORMAT, "layer '%s': LRN WITHIN_CHANNEL is not on the path", ly.name.c_str());
      }
    } else if (type == "Pooling") {
      ly.type = LayerType::Pooling;
      const Msg* pp = lm.msg("pooling_param");
      if (!pp || pp->str("pool", "MAX") != "MAX" || to_int(pp->str("kernel_size", "0"), "kernel_size") != 2 ||
          to_int(pp->str("stride", "1"), "stride") != 2 || to_int(pp->str("pad", "0"), "pad") != 0)
        fail(SIVO_EFORMAT, "layer '%s': only MAX 2x2 stride-2 pooling is on the path", ly.name.c_str());
      if (ly.tops.size() != 2) fail(SIVO_EFORMAT, "layer '%s': pooling must emit value and mask tops", ly.name.c_str());
    } else if (type == "Upsample") {
      ly.type = LayerType::Upsample;
      const Msg* up = lm.msg("upsample_param");
      int scale = up ? to_int(up->str("scale", "2"), "scale") : 2;
      if (scale != 2 || (up && (up->one("upsample_h") || up->one("scale_h") || up->one("pad_out_h"))))
        fail(SIVO_EFORMAT, "layer '%s': only scale-2 upsample is on the path", ly.name.c_str());
      if (ly.bottoms.size() != 2) fail(SIVO_EFORMAT, "layer '%s': upsample needs value and mask bottoms", ly.name.c_str());
    } else if (type == "Dropout") {
      ly.type = LayerType::Dropout;
      if (auto* dp = lm.msg("dropout_param")) {
        ly.dropout_ratio = to_float(dp->str("dropout_ratio", "0.5"), "dropout_ratio");
        ly.sample_weights_test = to_bool(dp->str("sample_weights_test", "false"));
      }
    } else if (type == "Softmax") {
      ly.type = LayerType::Softmax;
    } else {
      fail(SIVO_EFORMAT, "layer '%s': type '%s' is not on the SIVO perception path", ly.name.c_str(), type.c_str());
    }
    net.layers.push_back(std::move(ly));
  }
  if (net.layers.empty()) fail(SIVO_EFORMAT, "prototxt: no layers");
  return net;
}

NetSpec parse_prototxt_file(const std::string& path) {
  std::ifstream f(path);
  if (!f) fail(SIVO_ENOENT, "cannot open prototxt '%s'", path.c_str());
  std::stringstream ss;
  ss << f.rdbuf();
  return parse_prototxt_text(ss.str());
}

}  // namespace sivo
