/* TEST INFRASTRUCTURE ONLY - plain-C restatement of the reference's CPU aggregation loops.
 *
 * The checker for mid-sized parity tests (where the numpy restatement in nts_oracle.py is too
 * slow) and the fallback CPU baseline.  Never linked into or called from the product library.
 * Pinned against the golden vectors of the unmodified reference in tests/test_oracle_golden.py
 * (test_c_port_matches_golden).
 *
 * Each routine follows the reference's loop order:
 *   - forward:  for dst: for idx in CSC[dst]: out[dst] += in[src]*w   (core/ntsCPUFusedGraphOp.hpp:81-106,
 *               row primitive nts_comp core/ntsBaseOp.hpp:82-104: multiply, then add, float32)
 *   - backward: for src: for idx in CSR[src]: out[src] += in[dst]*w   (core/ntsCPUFusedGraphOp.hpp:123-143)
 * Both are the same segmented weighted gather-sum over (offsets, indices).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* out[r,:] += sum_{e in [offsets[r], offsets[r+1])} in[indices[e]-base, :] * (w ? w[e] : 1) */
void nts_oracle_segment_gather_sum(const uint32_t *offsets, const uint32_t *indices, const float *w,
                                   const float *in, float *out, uint32_t base, uint32_t n_rows,
                                   uint32_t feature_size) {
  const size_t F = feature_size;
#pragma omp parallel for schedule(dynamic, 64)
  for (long r = 0; r < (long)n_rows; r++) {
    float *o = out + (size_t)r * F;
    for (uint32_t e = offsets[r]; e < offsets[r + 1]; e++) {
      const float *x = in + (size_t)(indices[e] - base) * F;
      const float wt = w ? w[e] : 1.0f;
      for (size_t f = 0; f < F; f++)
        o[f] += x[f] * wt;
    }
  }
}

/* nts_norm_degree, core/ntsBaseOp.hpp:194-197 */
void nts_oracle_norm_degree(const uint32_t *src, const uint32_t *dst, const uint32_t *out_degree,
                            const uint32_t *in_degree, float *w, size_t n_edges) {
#pragma omp parallel for
  for (long e = 0; e < (long)n_edges; e++)
    w[e] = 1 / ((float)sqrt((double)out_degree[src[e]]) * (float)sqrt((double)in_degree[dst[e]]));
}

/* y[d,:] += sum_{e->d} msg[e,:]  (DistAggregateDst::forward, core/ntsDistCPUGraphOp.hpp:258-284) */
void nts_oracle_gather_msg_to_dst(const uint32_t *column_offset, const float *msg, float *y,
                                  uint32_t n_rows, uint32_t feature_size) {
  const size_t F = feature_size;
#pragma omp parallel for schedule(dynamic, 64)
  for (long d = 0; d < (long)n_rows; d++)
    for (uint32_t e = column_offset[d]; e < column_offset[d + 1]; e++)
      for (size_t f = 0; f < F; f++)
        y[(size_t)d * F + f] += msg[(size_t)e * F + f];
}

/* column-wise max-subtracted softmax over each destination segment
 * (DistEdgeSoftMax::forward, core/ntsDistCPUGraphOp.hpp:449-470 -> libtorch softmax(0)) */
void nts_oracle_edge_softmax(const uint32_t *column_offset, const float *m, float *a, uint32_t n_rows,
                             uint32_t cols) {
  const size_t H = cols;
#pragma omp parallel for schedule(dynamic, 64)
  for (long d = 0; d < (long)n_rows; d++) {
    uint32_t b = column_offset[d], e = column_offset[d + 1];
    for (size_t h = 0; h < H; h++) {
      if (e == b)
        continue;
      float mx = -INFINITY;
      for (uint32_t i = b; i < e; i++)
        mx = fmaxf(mx, m[(size_t)i * H + h]);
      float s = 0.f;
      for (uint32_t i = b; i < e; i++) {
        float v = expf(m[(size_t)i * H + h] - mx);
        a[(size_t)i * H + h] = v;
        s += v;
      }
      for (uint32_t i = b; i < e; i++)
        a[(size_t)i * H + h] /= s;
    }
  }
}
