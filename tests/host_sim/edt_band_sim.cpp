// edt_band_sim.cpp -- TEST INFRASTRUCTURE: runs the per-lane logic of the banded EDT kernels
// (sln_amodal_b200/csrc/edt_band.cuh, the very functions the CUDA kernels call) on the CPU, lane by lane,
// with the same workspace layout as sln_edt_sq, so that tests/test_edt_band_sim.py can hold the algorithm to
// the oracle without a GPU.  Nothing in the product links or calls this file.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../sln_amodal_b200/csrc/edt_band.cuh"

using namespace sln::edtband;

extern "C" int edt_band_sim(const uint8_t *map, int H, int W, int32_t *out, int *max_total)
{
    if (W % 32 != 0 || W > 1024 || H > 2048 || H < 1) return -1;
    const int nb = (H + BAND - 1) / BAND, tx = W / 32;
    const int cap = (H + W) * (H + W);
    std::vector<unsigned> flags((size_t)nb * tx, 0u), stk((size_t)nb * SLOTS * W, 0xdeadbeefu);
    std::vector<Words2> mwords((size_t)nb * W);
    int mt = 0;
    // ---- build pass: one "CTA" per band
    for (int b = 0; b < nb; ++b) {
        const int yb = b * BAND, rows = H - yb < BAND ? H - yb : BAND;
        static unsigned mask[BAND + 2][32];
        static int ldv[BAND][32], rdv[BAND][32];
        unsigned rowbits[BAND];
        for (int i = 0; i < BAND + 2; ++i) {
            const int yy = yb - 1 + i;
            for (int l = 0; l < 32; ++l) {
                unsigned z = 0u;
                if (yy >= 0 && yy < H && l < tx)
                    for (int k = 0; k < 32; ++k) z |= (unsigned)(map[(size_t)yy * W + l * 32 + k] == 0) << k;
                mask[i][l] = z;
            }
            if (i >= 1 && i <= BAND) {
                unsigned rb = 0u;
                if (yy < H)
                    for (int l = 0; l < tx; ++l) rb |= (unsigned)(mask[i][l] != FULLW) << l;
                rowbits[i - 1] = rb;
                int last = -NONE_D;
                for (int l = 0; l < 32; ++l) {
                    ldv[i - 1][l] = last == -NONE_D ? NONE_D : l * 32 - last;
                    const unsigned z = mask[i][l];
                    if (z) last = l * 32 + 31 - hd_clz(z);
                }
                int first = NONE_D;
                for (int l = 31; l >= 0; --l) {
                    rdv[i - 1][l] = first == NONE_D ? NONE_D : first - (l * 32 + 31);
                    const unsigned z = mask[i][l];
                    if (z) first = l * 32 + hd_ffs(z) - 1;
                }
            }
        }
        for (int s = 0; s < tx; ++s) {
            unsigned f = 0u;
            for (int r = 0; r < BAND; ++r) f |= ((rowbits[r] >> s) & 1u) << r;
            flags[(size_t)b * tx + s] = f;
            if (!f) continue;
            for (int lane = 0; lane < 32; ++lane) {
                const int x = s * 32 + lane;
                const bool above_exists = yb > 0, below_exists = yb + BAND < H;
                const bool az = above_exists && ((mask[0][s] >> lane) & 1u), af = above_exists && !az;
                const bool bz = below_exists && ((mask[BAND + 1][s] >> lane) & 1u), bf = below_exists && !bz;
                unsigned *sc = stk.data() + (size_t)b * SLOTS * W + x;
                const BuildResult res = band_build_lane(f, rows, yb, H, lane, az, af, bz, bf, sc, W,
                                                        [&](int r, unsigned &z, int &ld, int &rd) {
                                                            z = mask[r + 1][s];
                                                            ld = ldv[r][s];
                                                            rd = rdv[r][s];
                                                        });
                mwords[(size_t)b * W + x].x = res.fgw;
                mwords[(size_t)b * W + x].y = res.meta;
                if (meta_total(res.meta) > mt) mt = meta_total(res.meta);
                if (meta_total(res.meta) > SLOTS) return -2;
            }
        }
    }
    if (max_total) *max_total = mt;
    // ---- evaluation pass
    for (int b = 0; b < nb; ++b) {
        const int yb = b * BAND, rows = H - yb < BAND ? H - yb : BAND;
        for (int s = 0; s < tx; ++s) {
            if (!flags[(size_t)b * tx + s]) {
                for (int r = 0; r < rows; ++r) memset(out + (size_t)(yb + r) * W + s * 32, 0, 32 * sizeof(int32_t));
                continue;
            }
            for (int lane = 0; lane < 32; ++lane) {
                const int x = s * 32 + lane;
                const unsigned fgw = mwords[(size_t)b * W + x].x, mw = mwords[(size_t)b * W + x].y;
                // bands whose tile was never built hold no words: reading them would be a bug of the pruning logic
                int best[BAND];
                band_eval_lane(b, nb, yb, rows, cap, fgw, mw, stk.data() + (size_t)b * SLOTS * W + x, W, stk.data() + x,
                               (size_t)SLOTS * W, W, mwords.data() + x, W, best, 1,
                               [&](const unsigned *bp, int, int, bool, const unsigned *&ptr, int &stride) {
                                   ptr = bp;            // no staging on the host: the stack itself
                                   stride = W;
                               });
                for (int r = 0; r < rows; ++r) out[(size_t)(yb + r) * W + x] = ((fgw >> r) & 1u) ? best[r] : 0;
            }
        }
    }
    return 0;
}
