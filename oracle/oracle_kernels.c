/* ORACLE (test infrastructure, not product code).
 * fmaf dot products in the order CUDA kernel 3a (block_scores) uses: one fused-multiply-add chain per
 * (i, j) over d = 0..D-1 starting from +0.  Restates torch.bmm(q_pool, k_pool^T) of
 * /root/reference/rectified_spaattn/rectified_wan21_attn.py:203 and gapr_mask.py:26,32 in fp32. */
#include <math.h>

void oracle_dots(const float* x, const float* y, float* out, int nx, int ny, int d) {
    for (int i = 0; i < nx; ++i) {
        const float* xi = x + (long)i * d;
        for (int j = 0; j < ny; ++j) {
            const float* yj = y + (long)j * d;
            float acc = 0.0f;
            for (int t = 0; t < d; ++t) acc = fmaf(xi[t], yj[t], acc);
            out[(long)i * ny + j] = acc;
        }
    }
}
