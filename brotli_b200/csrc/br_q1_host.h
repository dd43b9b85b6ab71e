// br_q1_host.h -- host-visible interface of the quality-1 batch pipeline (br_q1.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>
struct BrQ1Job;
struct BrQ1Stats {
  float ms_total, ms_h2d, ms_parse, ms_code, ms_pack, ms_d2h;
  uint64_t streams, fragments, blocks, in_bytes, out_bytes, launches;
};
// Device-resident batch: the streams sit back to back (any layout) in ONE device buffer, stream s at d_in + in_off[s];
// the compressed streams are packed densely (16-byte aligned starts) into d_out, stream s at d_out + out_off[s].
struct BrQ1Packed {
  const uint8_t* d_in; const uint64_t* in_off;   // in_off: host array [count]
  uint8_t* d_out; size_t out_cap; uint64_t* out_off;   // out_off: host array [count + 1] (last = total dense bytes)
};
extern "C" {
BrQ1Job* br_q1_job_create(void);
void br_q1_job_destroy(BrQ1Job*);
const BrQ1Stats* br_q1_job_stats(const BrQ1Job*);
// Compresses `count` independent streams at quality 1.  calls/ncalls (nullable): the sizes of the
// CompressStream calls that delivered each stream (the reference cuts fragments per call); null = one
// call.  out_n: capacity in, size out.  ok[s] = 0 when out[s] was too small.  packed (nullable): the
// device-resident form above (in[] / out[] are then unused, out_n receives the sizes).  with_header / end_op: 1 / 2 for whole streams; a segment of a stream that
// is cut by FLUSH calls has no header after the first one (0) and ends with byte padding (end_op 1).
// end_op 0: a segment with more input behind it (may end mid-byte: end_bit[s] = its bit length, byte 0 leaves start_bits[s]
// low bits free for the previous segment's tail).  Returns 1 when every stream was compressed.
int br_q1_compress_batch(BrQ1Job* job, int lgwin, size_t count, const uint8_t* const* in, const size_t* in_n,
                         const size_t* const* calls, const size_t* ncalls, const BrQ1Packed* packed,
                         uint8_t* const* out, size_t* out_n, int* ok, int threads, int with_header, int end_op,
                         const uint32_t* start_bits, uint32_t* end_bit);
}
