// Test harness (CPU): the host mirror's sink (reconstruction_b200/host/CCloudOptimization.cpp) linked against a STUB of the three
// C-ABI entry points it calls, so the hand-over logic around the GPU filter - pair ranges, records filtered ahead of filter(),
// the one-off reservation, the background writer of tmp/cloud_filter.ply (latest state wins, flushed by run()) - runs without a
// GPU.  The stub "filter" keeps every point whose index is not a multiple of 3 and emits (x, y, z, pair, index, n, 1).
// usage: sink_handover <dir> <pairs> <points of pair 0> <points of pair 1> ...
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../include/stereo_b200.h"
#include "../../reconstruction_b200/host/CCloudOptimization.h"

static int g_calls = 0;
extern "C" {
const char* sb200_status_string(int) { return "stub"; }
const char* sb200_sink_last_error(void) { return ""; }
int sb200_sink_filter(int device, const double* xyz, int64_t n, int, double, double, const double*, float* out, int32_t*, int64_t capacity,
                      int64_t* n_kept, double* stats5) {
  g_calls++;
  int64_t k = 0;
  for (int64_t i = 0; i < n; i++) {
    if (i % 3 == 0) continue;
    if (k >= capacity) return SB200_ERR_BAD_ARG;
    float* r = out + 7 * k++;
    r[0] = (float)xyz[3 * i]; r[1] = (float)xyz[3 * i + 1]; r[2] = (float)xyz[3 * i + 2];
    r[3] = (float)device; r[4] = (float)i; r[5] = (float)n; r[6] = 1.0f;
  }
  *n_kept = k;
  if (stats5) { stats5[0] = 1; stats5[1] = 2; stats5[2] = 3; stats5[3] = 0.5; stats5[4] = 0; }
  std::this_thread::sleep_for(std::chrono::milliseconds(2));
  return SB200_OK;
}
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  if (chdir(argv[1]) != 0) return 2;
  const int P = atoi(argv[2]);
  if (argc < 3 + P) return 2;
  CManageData data;
  data.m_CampairNum = P;
  data.cam.assign(P, std::vector<camera>(2));
  data.outfilename = std::string(argv[1]) + "/out.ply";
  {
    CCloudOptimization sink;
    sink.Init(100, 1, 50, 2, 2.5, &data, false);
    std::vector<std::vector<double>> pts(P);
    size_t total = 0;
    for (int p = 0; p < P; p++) {
      const size_t n = (size_t)atol(argv[3 + p]);
      pts[p].resize(3 * n);
      for (size_t i = 0; i < n; i++) { pts[p][3 * i] = p; pts[p][3 * i + 1] = (double)i; pts[p][3 * i + 2] = 0.25 * (double)i + p; }
      total += n;
    }
    sink.Reserve(total);
    const double* base = sink.xyz.capacity() >= 3 * total ? sink.xyz.data() : nullptr;
    // odd pairs are filtered ahead by "workers" (device 7), as the matcher does; even pairs inside filter() (the sink's device 0)
    std::vector<std::thread> workers;
    for (int p = 1; p < P; p += 2)
      workers.emplace_back([&, p]() {
        SinkRecords rec;
        size_t kept = 0;
        double st[5];
        std::string err;
        if (pts[p].empty()) return;
        if (sink.FilterPoints(p, 7, pts[p].data(), pts[p].size() / 3, rec, kept, st, err)) sink.StoreFiltered(p, std::move(rec), kept, st);
      });
    for (auto& w : workers) w.join();
    for (int p = 0; p < P; p++) {
      std::vector<unsigned char> bgr(pts[p].size(), (unsigned char)(10 + p));
      sink.InsertPoints(pts[p].data(), bgr.data(), pts[p].size() / 3);
      sink.filter(p);
    }
    if (base && sink.xyz.data() != base) { printf("the merged buffer moved after Reserve\n"); return 1; }
    sink.run();
    size_t kept = 0;
    for (size_t k : sink.kept_per_pair) kept += k;
    printf("pairs %d points %zu kept %zu filter_calls %d\n", P, sink.xyz.size() / 3, kept, g_calls);
  }  // the destructor joins the writer
  return 0;
}
