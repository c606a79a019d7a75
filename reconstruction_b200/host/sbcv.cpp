#include "sbcv.h"

#include <ctype.h>
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

bool FileStorage::open(const std::string& path) {
  opened_ = false;
  nodes_.clear();
  std::ifstream in(path.c_str());
  if (!in) { err_ = "cannot open " + path; return false; }
  std::vector<std::string> lines;
  for (std::string l; std::getline(in, l);) {
    if (!l.empty() && l.back() == '\r') l.pop_back();
    lines.push_back(l);
  }
  size_t i = 0;
  auto blank = [&](const std::string& l) { const std::string t = trim(l); return t.empty() || t[0] == '#'; };
  // gathers "[ ... ]" possibly spread over several lines, starting with `first` (text after the key)
  auto gather_flow = [&](std::string first, size_t& idx) {
    std::string body = first;
    while (body.find(']') == std::string::npos && idx < lines.size()) body += " " + trim(lines[idx++]);
    const size_t a = body.find('['), b = body.rfind(']');
    return (a == std::string::npos || b == std::string::npos || b < a) ? std::string() : body.substr(a + 1, b - a - 1);
  };
  while (i < lines.size()) {
    const std::string& l = lines[i];
    if (blank(l) || l[0] == '%' || trim(l) == "---" || trim(l) == "...") { i++; continue; }
    if (indent_of(l) != 0) { i++; continue; }  // stray continuation
    const size_t colon = l.find(':');
    if (colon == std::string::npos) { i++; continue; }
    const std::string key = trim(l.substr(0, colon));
    std::string rest = trim(l.substr(colon + 1));
    i++;
    FileNode n;
    if (rest.compare(0, 15, "!!opencv-matrix") == 0) {
      n.kind = FileNode::MATRIX;
      while (i < lines.size() && (blank(lines[i]) || indent_of(lines[i]) > 0)) {
        if (blank(lines[i])) { i++; continue; }
        const std::string t = trim(lines[i]);
        const size_t c2 = t.find(':');
        i++;
        if (c2 == std::string::npos) continue;
        const std::string k2 = trim(t.substr(0, c2)), v2 = trim(t.substr(c2 + 1));
        if (k2 == "rows") n.rows = atoi(v2.c_str());
        else if (k2 == "cols") n.cols = atoi(v2.c_str());
        else if (k2 == "dt") n.dt = unquote(v2);
        else if (k2 == "data") {
          for (const std::string& s : split_flow(gather_flow(v2, i))) n.values.push_back(strtod(s.c_str(), nullptr));
        }
      }
      if ((size_t)n.rows * n.cols != n.values.size() && !n.values.empty() && n.rows > 0 && n.cols > 0 &&
          n.values.size() % ((size_t)n.rows * n.cols) != 0) {
        err_ = "matrix " + key + ": rows*cols does not match data";
        return false;
      }
    } else if (!rest.empty() && rest[0] == '[') {
      n.kind = FileNode::SEQ;
      n.seq = split_flow(gather_flow(rest, i));
    } else if (rest.empty()) {
      n.kind = FileNode::SEQ;
      while (i < lines.size() && (blank(lines[i]) || indent_of(lines[i]) > 0 || trim(lines[i])[0] == '-')) {
        if (blank(lines[i])) { i++; continue; }
        std::string t = trim(lines[i]);
        if (t[0] != '-') break;
        i++;
        n.seq.push_back(unquote(t.substr(1)));
      }
    } else {
      n.kind = FileNode::SCALAR;
      n.scalar = unquote(rest);
    }
    nodes_[key] = n;
  }
  opened_ = true;
  return true;
}

const FileNode& FileStorage::operator[](const std::string& key) const {
  static const FileNode none;
  auto it = nodes_.find(key);
  return it == nodes_.end() ? none : it->second;
}

std::vector<std::string> FileStorage::keys() const {
  std::vector<std::string> k;
  for (auto& kv : nodes_) k.push_back(kv.first);
  return k;
}

void operator>>(const FileNode& n, int& v) { v = n.kind == FileNode::SCALAR ? (int)strtol(n.scalar.c_str(), nullptr, 10) : 0; }
void operator>>(const FileNode& n, double& v) { v = n.kind == FileNode::SCALAR ? strtod(n.scalar.c_str(), nullptr) : 0.0; }
void operator>>(const FileNode& n, std::string& v) { v = n.kind == FileNode::SCALAR ? n.scalar : std::string(); }
void operator>>(const FileNode& n, std::vector<std::string>& v) {
  v.clear();
  if (n.kind == FileNode::SEQ) v = n.seq;
  else if (n.kind == FileNode::SCALAR) v.push_back(n.scalar);
}
void operator>>(const FileNode& n, Mat& m) {
  m.release();
  if (n.kind != FileNode::MATRIX || n.rows <= 0 || n.cols <= 0 || n.values.size() < (size_t)n.rows * n.cols) return;
  const bool u8 = n.dt == "u";
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
