// br_q1_host.h -- host-visible interface of the quality-1 batch pipeline (br_q1.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>
struct BrQ1Job;
struct BrQ1Stats {
  float ms_total, ms_h2d, ms_parse, ms_code, ms_pack, ms_d2h;
  uint64_t streams, fragments, blocks, in_bytes, out_bytes, launches;
};
extern "C" {
BrQ1Job* br_q1_job_create(void);
void br_q1_job_destroy(BrQ1Job*);
const BrQ1Stats* br_q1_job_stats(const BrQ1Job*);
// Compresses `count` independent streams at quality 1.  calls/ncalls (nullable): the sizes of the
// CompressStream calls that delivered each stream (the reference cuts fragments per call); null = one
// call.  out_n: capacity in, size out.  ok[s] = 0 when out[s] was too small.  inputs_on_device: in[] and
// out[] are device pointers.  with_header / end_op: 1 / 2 for whole streams; a segment of a stream that
// is cut by FLUSH calls has no header after the first one (0) and ends with byte padding (end_op 1).
// Returns 1 when every stream was compressed.
int br_q1_compress_batch(BrQ1Job* job, int lgwin, size_t count, const uint8_t* const* in, const size_t* in_n,
                         const size_t* const* calls, const size_t* ncalls, int inputs_on_device,
                         uint8_t* const* out, size_t* out_n, int* ok, int threads, int with_header, int end_op);
}
