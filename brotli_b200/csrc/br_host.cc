// br_host.cc -- host-side tables: the embedded format tables blob and the log2 table that
// makes the device's double-precision entropy decisions agree with the reference's
// FastLog2 (c/enc/fast_log.h:51) on THIS host: entries < 256 are float-rounded exactly like
// the literals of c/enc/fast_log.c:13, the rest are this machine's libm log2() -- the same
// function the reference encoder would call here.
#include <math.h>
#include <stdint.h>
#include <mutex>
#include <thread>
#include <vector>

extern "C" {
extern const unsigned char br_tables_blob[] = {
#include "br_tables_blob.inc"
};
extern const unsigned int br_tables_blob_len = sizeof(br_tables_blob);
}

static std::vector<double> g_log2;
static std::once_flag g_log2_once;

extern "C" const double* br_host_log2_table(uint32_t* n) {
  std::call_once(g_log2_once, [] {
    const uint32_t N = (1u << 24) + 2;   // metablocks are at most 1 << 24 bytes (quality.h:103)
    g_log2.resize(N);
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 16) nt = 16;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([=] {
        for (uint32_t i = t; i < N; i += nt) {
          if (i == 0) g_log2[i] = 0.0;
          else if (i < 256) g_log2[i] = (double)(float)log2((double)i);
          else g_log2[i] = log2((double)i);
        }
      });
    for (auto& x : th) x.join();
  });
  *n = (uint32_t)g_log2.size();
  return g_log2.data();
}
