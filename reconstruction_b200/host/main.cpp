// reconstruction <config.yml>   — CLI of the reference (reconstruction/main.cpp:5-24) on the B200 path.
//   --dump-config <config.yml>   parse only (no GPU): print what CManageData::Init read, as JSON
//   --decode <in> <out.pnm> [gray]   decode one image file with the native readers (no GPU) and write it as PNM
//   --dump-yaml <file.yml>       parse only: print every node the OpenCV-YAML reader found, as JSON
//   --decode-bench <in> <reps>   time the native decoder on one file (colour), print JSON
//   --prefetch-test <threads> <files...>   decode every file (colour and grey, each requested twice) through the
//                                background ImagePrefetcher; print "<file> <mode> <cols> <rows> <fnv1a>" per consumer
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <utility>
#include <vector>

#include "CReconstruction.h"

static void dump(const CManageData& d) {
  printf("{\"filepath\": \"%s\", \"outfilename\": \"%s\", \"isoutput\": %d, \"PyrmNum\": %d, \"LowestLevelWidth\": %d, "
         "\"LowestLevelHeight\": %d, \"OriginWidth\": %d, \"OriginHeight\": %d, \"CameraNum\": %d, \"pairs\": [",
         d.m_FilePath.c_str(), d.outfilename.c_str(), d.isoutput, d.m_PyrmNum, d.m_LowestLevelSize.width, d.m_LowestLevelSize.height,
         d.m_OriginSize.width, d.m_OriginSize.height, d.m_CameraNum);
  for (int i = 0; i < d.m_CampairNum; i++) {
    printf("%s{", i ? ", " : "");
    for (int k = 0; k < 2; k++) {
      const camera& c = d.cam[i][k];
      printf("%s\"cam%d\": {\"id\": %d, \"image\": \"%s\", \"mask\": \"%s\", \"K\": [", k ? ", " : "", k, c.camID, c.image_name.c_str(),
             c.mask_name.c_str());
      for (int r = 0; r < 9; r++) printf("%s%.17g", r ? ", " : "", c.MatIntrinsics.ptr<double>()[r]);
      printf("], \"Rt\": [");
      for (int r = 0; r < 12; r++) printf("%s%.17g", r ? ", " : "", c.MatExtrinsics.ptr<double>()[r]);
      printf("], \"center\": [%.17g, %.17g, %.17g]}", c.CamCenter.ptr<double>()[0], c.CamCenter.ptr<double>()[1], c.CamCenter.ptr<double>()[2]);
    }
    printf("}");
  }
  printf("]}\n");
}

static void json_string(const std::string& s) {
  putchar('"');
  for (unsigned char ch : s) {
    if (ch == '"' || ch == '\\') printf("\\%c", ch);
    else if (ch < 0x20) printf("\\u%04x", ch);
    else putchar(ch);
  }
  putchar('"');
}

static void dump_node(const sbcv::FileNode& n) {
  switch (n.kind) {
    case sbcv::FileNode::SCALAR: json_string(n.scalar); break;
    case sbcv::FileNode::SEQ:
      printf("[");
      for (size_t i = 0; i < n.seq.size(); i++) { if (i) printf(", "); json_string(n.seq[i]); }
      printf("]");
      break;
    case sbcv::FileNode::MATRIX:
      printf("{\"rows\": %d, \"cols\": %d, \"channels\": %d, \"dt\": \"%s\", \"hex\": [", n.rows, n.cols, n.channels, n.dt.c_str());
      for (size_t i = 0; i < n.values.size(); i++) {  // exact: the IEEE bits of every element
        unsigned long long u;
        memcpy(&u, &n.values[i], 8);
        printf("%s\"%016llx\"", i ? ", " : "", u);
      }
      printf("]}");
      break;
    case sbcv::FileNode::MAP:
      printf("{");
      for (size_t i = 0; i < n.child_keys.size(); i++) {
        if (i) printf(", ");
        json_string(n.child_keys[i]);
        printf(": ");
        dump_node(n.child_nodes[i]);
      }
      printf("}");
      break;
    default: printf("null");
  }
}

int main(int Argc, char** Argv) {
  const auto start = std::chrono::steady_clock::now();
  auto elapsed = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count(); };
  CReconstrction recon;
  if (Argc <= 1) {
    printf("USAGE: reconstruction your_config_file.yml\n");
    return -1;
  }
  if (Argc >= 4 && strcmp(Argv[1], "--decode") == 0) {
    sbcv::Mat img;
    if (!sbcv::imread(Argv[2], img, Argc >= 5 && strcmp(Argv[4], "gray") == 0)) {
      printf("read image %s error: %s\n", Argv[2], sbcv::imread_error().c_str());
      return 1;
    }
    return sbcv::imwrite_pnm(Argv[3], img) ? 0 : 1;
  }
  if (Argc >= 4 && strcmp(Argv[1], "--prefetch-test") == 0) {
    std::vector<std::pair<std::string, bool>> req;
    for (int rep = 0; rep < 2; rep++)
      for (int i = 3; i < Argc; i++) {
        req.emplace_back(Argv[i], false);
        req.emplace_back(Argv[i], true);
      }
    sbcv::ImagePrefetcher pf;
    pf.start(req, atoi(Argv[2]));
    int bad = 0;
    for (const auto& r : req) {
      sbcv::Mat m;
      std::string err;
      if (!pf.get(r.first, r.second, m, &err)) {
        printf("%s %s error %s\n", r.first.c_str(), r.second ? "gray" : "color", err.c_str());
        bad++;
        continue;
      }
      unsigned long long h = 1469598103934665603ull;
      for (size_t k = 0; k < m.total_bytes(); k++) h = (h ^ m.data[k]) * 1099511628211ull;
      printf("%s %s %d %d %016llx\n", r.first.c_str(), r.second ? "gray" : "color", m.cols, m.rows, h);
    }
    sbcv::Mat extra;
    std::string err;
    if (Argc > 3 && pf.get(Argv[3], false, extra, &err) && !extra.empty()) printf("%s color served after its last consumer\n", Argv[3]);
    return bad ? 1 : 0;
  }
  if (Argc >= 4 && strcmp(Argv[1], "--decode-bench") == 0) {
    FILE* fp = fopen(Argv[2], "rb");
    if (!fp) { printf("cannot open %s\n", Argv[2]); return 1; }
    std::string bytes;
    char buf[1 << 16];
    for (size_t n; (n = fread(buf, 1, sizeof buf, fp)) > 0;) bytes.append(buf, n);
    fclose(fp);
    const int reps = atoi(Argv[3]) > 0 ? atoi(Argv[3]) : 1;
    sbcv::Mat img;
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
      const auto t0 = std::chrono::steady_clock::now();
      if (!sbcv::imdecode(reinterpret_cast<const uint8_t*>(bytes.data()), bytes.size(), img, false)) {
        printf("decode error: %s\n", sbcv::imread_error().c_str());
        return 1;
      }
      best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    printf("{\"width\": %d, \"height\": %d, \"file_bytes\": %zu, \"best_ms\": %.3f, \"Mpix_per_s\": %.2f}\n", img.cols, img.rows, bytes.size(),
           1e3 * best, img.cols * (double)img.rows / best / 1e6);
    return 0;
  }
  if (Argc >= 3 && strcmp(Argv[1], "--dump-yaml") == 0) {
    sbcv::FileStorage fs(Argv[2], sbcv::FileStorage::READ);
    if (!fs.isOpened()) {
      printf("cannot open file %s: %s\n", Argv[2], fs.error().c_str());
      return 1;
    }
    printf("{");
    bool first = true;
    for (const std::string& k : fs.keys()) {
      if (!first) printf(", ");
      first = false;
      json_string(k);
      printf(": ");
      dump_node(fs[k]);
    }
    printf("}\n");
    return 0;
  }
  if (Argc >= 3 && strcmp(Argv[1], "--dump-config") == 0) {
    if (recon.Init(Argv[2]) == false) return -1;
    dump(recon.m_ImageData);
    return 0;
  }
  if (recon.Init(Argv[1]) == false) return -1;
  recon.m_Matching.MatchAllLayer();
  printf("Matching time: %.3f s (GPU path %.3f s)\n", elapsed(), recon.m_Matching.gpu_seconds);
  if (recon.m_Matching.last_status != 0) {
    printf("matching failed: %s\n", recon.m_Matching.last_error.c_str());
    return 1;
  }
  recon.m_CloudOptimization.run();
  printf("total time: %.3f s\n", elapsed());
  return 0;
}
