#include "sbcv.h"

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <sstream>

namespace sbcv {

void Mat::create(int r, int c, int type) {
  rows = r; cols = c; type_ = type;
  store_ = std::make_shared<std::vector<uint8_t>>((size_t)r * c * elemSize());
  data = store_->data();
}

Mat Mat::clone() const {
  Mat m;
  if (empty()) return m;
  m.create(rows, cols, type_);
  memcpy(m.data, data, total_bytes());
  return m;
}

// ---------------------------------------------------------------------------------------- YAML
static std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) a++;
  while (b > a && isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}

static std::string unquote(const std::string& t) {
  std::string s = trim(t);
  if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) {
    std::string o;
    for (size_t i = 1; i + 1 < s.size(); i++) {
      if (s[i] == '\\' && s.front() == '"' && i + 2 < s.size()) {  // OpenCV escapes \\ and \" inside double quotes
        i++;
        o += s[i] == 'n' ? '\n' : s[i] == 't' ? '\t' : s[i];
      } else {
        o += s[i];
      }
    }
    return o;
  }
  return s;
}

// split a flow sequence body "a, "b c", 3" on commas outside quotes
static std::vector<std::string> split_flow(const std::string& body) {
  std::vector<std::string> out;
  std::string cur;
  char q = 0;
  for (char ch : body) {
    if (q) {
      cur += ch;
      if (ch == q) q = 0;
    } else if (ch == '"' || ch == '\'') {
      q = ch;
      cur += ch;
    } else if (ch == ',') {
      if (!trim(cur).empty()) out.push_back(unquote(cur));
      cur.clear();
    } else {
      cur += ch;
    }
  }
  if (!trim(cur).empty()) out.push_back(unquote(cur));
  return out;
}

static int indent_of(const std::string& l) {
  int n = 0;
  while (n < (int)l.size() && l[n] == ' ') n++;
  return n;
}

static double yaml_number(const std::string& t) {
  const std::string s = trim(t);
  const char* c = s.c_str();
  const bool neg = *c == '-';
  if (*c == '-' || *c == '+') c++;
  if (*c == '.' && (c[1] == 'I' || c[1] == 'i')) return neg ? -HUGE_VAL : HUGE_VAL;  // .Inf
  if (*c == '.' && (c[1] == 'N' || c[1] == 'n')) return NAN;                             // .Nan
  return strtod(s.c_str(), nullptr);
}

namespace {
struct YamlParser {
  std::vector<std::string> lines;
  std::string err;
  static bool blank(const std::string& l) { const std::string t = trim(l); return t.empty() || t[0] == '#'; }
  // gathers "[ ... ]" possibly spread over several lines, starting with `first` (text after the key)
  std::string gather_flow(const std::string& first, size_t& idx) const {
    std::string body = first;
    while (body.find(']') == std::string::npos && idx < lines.size()) body += " " + trim(lines[idx++]);
    const size_t a = body.find('['), b = body.rfind(']');
    return (a == std::string::npos || b == std::string::npos || b < a) ? std::string() : body.substr(a + 1, b - a - 1);
  }
  // first ':' that ends a key: outside quotes and followed by a blank or the end of the line
  static size_t key_colon(const std::string& l) {
    char q = 0;
    for (size_t k = 0; k < l.size(); k++) {
      const char ch = l[k];
      if (q) { if (ch == q) q = 0; continue; }
      if (ch == '"' || ch == '\'') { q = ch; continue; }
      if (ch == ':' && (k + 1 == l.size() || l[k + 1] == ' ' || l[k + 1] == '\t')) return k;
    }
    return std::string::npos;
  }
  size_t next_content(size_t i) const {
    while (i < lines.size() && blank(lines[i])) i++;
    return i;
  }
  // mapping whose keys sit at column `indent`; stops at the first line indented less
  bool parse_map(size_t& i, int indent, FileNode& out) {
    out.kind = FileNode::MAP;
    while (i < lines.size()) {
      const std::string& l = lines[i];
      const std::string t = trim(l);
      if (blank(l) || (indent == 0 && (l[0] == '%' || t == "---" || t == "..."))) { i++; continue; }
      const int ind = indent_of(l);
      if (ind < indent) return true;
      if (ind > indent) { i++; continue; }  // stray continuation
      const size_t colon = key_colon(l);
      if (colon == std::string::npos) { i++; continue; }
      const std::string key = unquote(l.substr(0, colon));
      const std::string rest = trim(l.substr(colon + 1));
      i++;
      FileNode n;
      if (rest.compare(0, 15, "!!opencv-matrix") == 0) {
        n.kind = FileNode::MATRIX;
        while (i < lines.size() && (blank(lines[i]) || indent_of(lines[i]) > indent)) {
          if (blank(lines[i])) { i++; continue; }
          const std::string m = trim(lines[i]);
          const size_t c2 = m.find(':');
          i++;
          if (c2 == std::string::npos) continue;
          const std::string k2 = trim(m.substr(0, c2)), v2 = trim(m.substr(c2 + 1));
          if (k2 == "rows") n.rows = atoi(v2.c_str());
          else if (k2 == "cols") n.cols = atoi(v2.c_str());
          else if (k2 == "dt") {
            n.dt = unquote(v2);
            n.channels = isdigit((unsigned char)n.dt[0]) ? atoi(n.dt.c_str()) : 1;
            if (n.channels < 1) n.channels = 1;
            while (!n.dt.empty() && isdigit((unsigned char)n.dt[0])) n.dt.erase(0, 1);
          } else if (k2 == "data") {
            for (const std::string& s : split_flow(gather_flow(v2, i))) n.values.push_back(yaml_number(s));
          }
        }
        if (n.rows < 0 || n.cols < 0 || n.values.size() != (size_t)n.rows * n.cols * n.channels) {
          err = "matrix " + key + ": rows*cols does not match data";
          return false;
        }
      } else if (!rest.empty() && rest[0] == '[') {
        n.kind = FileNode::SEQ;
        n.seq = split_flow(gather_flow(rest, i));
      } else if (rest.empty() || rest[0] == '#') {
        const size_t j = next_content(i);
        if (j < lines.size() && trim(lines[j])[0] == '-' && indent_of(lines[j]) >= indent) {  // block sequence
          n.kind = FileNode::SEQ;
          while (i < lines.size()) {
            if (blank(lines[i])) { i++; continue; }
            const std::string m = trim(lines[i]);
            if (m[0] != '-' || indent_of(lines[i]) < indent) break;
            i++;
            n.seq.push_back(unquote(m.substr(1)));
          }
        } else if (j < lines.size() && indent_of(lines[j]) > indent) {  // nested mapping
          i = j;
          if (!parse_map(i, indent_of(lines[j]), n)) return false;
        } else {
          n.kind = FileNode::SEQ;  // "key:" with nothing below: an empty sequence
        }
      } else {
        n.kind = FileNode::SCALAR;
        n.scalar = unquote(rest);
      }
      bool replaced = false;
      for (size_t k = 0; k < out.child_keys.size(); k++)
        if (out.child_keys[k] == key) { out.child_nodes[k] = n; replaced = true; }
      if (!replaced) { out.child_keys.push_back(key); out.child_nodes.push_back(n); }
    }
    return true;
  }
};
}  // namespace

bool FileStorage::open(const std::string& path) {
  opened_ = false;
  root_ = FileNode();
  std::ifstream in(path.c_str());
  if (!in) { err_ = "cannot open " + path; return false; }
  YamlParser yp;
  for (std::string l; std::getline(in, l);) {
    if (!l.empty() && l.back() == '\r') l.pop_back();
    if (yp.lines.empty() && l.size() >= 3 && (unsigned char)l[0] == 0xEF && (unsigned char)l[1] == 0xBB && (unsigned char)l[2] == 0xBF) l.erase(0, 3);
    yp.lines.push_back(l);
  }
  size_t i = 0;
  if (!yp.parse_map(i, 0, root_)) { err_ = yp.err; return false; }
  opened_ = true;
  return true;
}

const FileNode& FileNode::operator[](const std::string& key) const {
  static const FileNode none;
  for (size_t k = 0; k < child_keys.size(); k++)
    if (child_keys[k] == key) return child_nodes[k];
  return none;
}

const FileNode& FileStorage::operator[](const std::string& key) const { return root_[key]; }

std::vector<std::string> FileStorage::keys() const { return root_.child_keys; }

void operator>>(const FileNode& n, int& v) {  // cv::FileNode rounds a real to the nearest integer
  if (n.kind != FileNode::SCALAR) { v = 0; return; }
  char* e = nullptr;
  const long l = strtol(n.scalar.c_str(), &e, 10);
  v = (e && *e == 0) ? (int)l : (int)lrint(yaml_number(n.scalar));
}
void operator>>(const FileNode& n, double& v) { v = n.kind == FileNode::SCALAR ? yaml_number(n.scalar) : 0.0; }
void operator>>(const FileNode& n, std::string& v) { v = n.kind == FileNode::SCALAR ? n.scalar : std::string(); }
void operator>>(const FileNode& n, std::vector<std::string>& v) {
  v.clear();
  if (n.kind == FileNode::SEQ) v = n.seq;
  else if (n.kind == FileNode::SCALAR) v.push_back(n.scalar);
}
void operator>>(const FileNode& n, Mat& m) {
  m.release();
  if (n.kind != FileNode::MATRIX || n.rows <= 0 || n.cols <= 0 || n.values.size() < (size_t)n.rows * n.cols * n.channels) return;
  const bool u8 = n.dt == "u";
  if (n.channels == 3 && u8) {
    m.create(n.rows, n.cols, SB_8UC3);
    for (size_t k = 0; k < n.values.size(); k++) m.data[k] = (uint8_t)n.values[k];
    return;
  }
  if (n.channels != 1) return;  // no multi-channel real matrices in this mirror
  m.create(n.rows, n.cols, u8 ? SB_8UC1 : SB_64FC1);
  for (int r = 0; r < n.rows; r++)
    for (int c = 0; c < n.cols; c++) {
      const double v = n.values[(size_t)r * n.cols + c];
      if (u8) m.at<uint8_t>(r, c) = (uint8_t)v; else m.at<double>(r, c) = v;
    }
}

// ---------------------------------------------------------------------------------------- PNM
static bool pnm_token(FILE* fp, int& v) {
  int c = fgetc(fp);
  for (;;) {
    while (c != EOF && isspace(c)) c = fgetc(fp);
    if (c == '#') { while (c != EOF && c != '\n') c = fgetc(fp); continue; }
    break;
  }
  if (c == EOF || !isdigit(c)) return false;
  v = 0;
  while (c != EOF && isdigit(c)) { v = v * 10 + (c - '0'); c = fgetc(fp); }
  return true;  // the single whitespace after the token has been consumed
}

bool imread_pnm(const std::string& path, Mat& out, bool grayscale) {
  out.release();
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) return false;
  char magic[3] = {0, 0, 0};
  if (fread(magic, 1, 2, fp) != 2 || magic[0] != 'P' || (magic[1] != '5' && magic[1] != '6')) { fclose(fp); return false; }
  int w = 0, h = 0, mx = 0;
  if (!pnm_token(fp, w) || !pnm_token(fp, h) || !pnm_token(fp, mx) || w <= 0 || h <= 0 || mx != 255) { fclose(fp); return false; }
  const int cn = magic[1] == '6' ? 3 : 1;
  std::vector<uint8_t> buf((size_t)w * h * cn);
  const bool ok = fread(buf.data(), 1, buf.size(), fp) == buf.size();
  fclose(fp);
  if (!ok) return false;
  if (grayscale) {
    out.create(h, w, SB_8UC1);
    if (cn == 1) memcpy(out.data, buf.data(), buf.size());
    else
      for (size_t i = 0; i < (size_t)w * h; i++) {  // fixed-point BT.601 as cv::cvtColor (R 4899, G 9617, B 1868, >> 14)
        const int r = buf[3 * i], g = buf[3 * i + 1], b = buf[3 * i + 2];
        out.data[i] = (uint8_t)((r * 4899 + g * 9617 + b * 1868 + (1 << 13)) >> 14);
      }
  } else {
    out.create(h, w, SB_8UC3);
    for (size_t i = 0; i < (size_t)w * h; i++) {
      if (cn == 3) { out.data[3 * i] = buf[3 * i + 2]; out.data[3 * i + 1] = buf[3 * i + 1]; out.data[3 * i + 2] = buf[3 * i]; }
      else out.data[3 * i] = out.data[3 * i + 1] = out.data[3 * i + 2] = buf[i];
    }
  }
  return true;
}

bool imwrite_pnm(const std::string& path, const Mat& m) {
  if (m.empty() || (m.type() != SB_8UC1 && m.type() != SB_8UC3)) return false;
  FILE* fp = fopen(path.c_str(), "wb");
  if (!fp) return false;
  fprintf(fp, "P%c\n%d %d\n255\n", m.type() == SB_8UC3 ? '6' : '5', m.cols, m.rows);
  if (m.type() == SB_8UC1) fwrite(m.data, 1, m.total_bytes(), fp);
  else {
    std::vector<uint8_t> row((size_t)m.cols * 3);
    for (int y = 0; y < m.rows; y++) {
      const uint8_t* p = m.ptr<uint8_t>(y);
      for (int x = 0; x < m.cols; x++) { row[3 * x] = p[3 * x + 2]; row[3 * x + 1] = p[3 * x + 1]; row[3 * x + 2] = p[3 * x]; }
      fwrite(row.data(), 1, row.size(), fp);
    }
  }
  fclose(fp);
  return true;
}

}  // namespace sbcv
