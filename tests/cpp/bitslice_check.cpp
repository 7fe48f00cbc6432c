// Host check of the bit-sliced primitives of hyslam_b200/csrc/fast_bitslice.cuh against scalar restatements:
// bs_transpose + bs_corners vs a per-pixel FAST-9/16 segment test.
// Built and run by tests/test_bitslice_host.py (g++ only, no GPU).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "../../hyslam_b200/csrc/fast_bitslice.cuh"

static const int NSEG = 3, W = 32 * NSEG, H = 24, PLP = 8 * (NSEG + 2) + 4;
static uint32_t rng_state = 12345u;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

static void to_planes(const uint8_t *img, uint32_t *planes)
{
    memset(planes, 0, sizeof(uint32_t) * H * PLP);
    for (int r = 0; r < H; r++)
        for (int s = 0; s < NSEG; s++) {
            uint32_t w[8], P[8];
            memcpy(w, img + r * W + 32 * s, 32);
            bs_transpose(w, P);
            memcpy(planes + r * PLP + (s + 1) * 8, P, 32);
        }
}

int main()
{
    static uint8_t img[H * W];
    static uint32_t planes[H * PLP];
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    long bad = 0, corners = 0;
    for (int trial = 0; trial < 200; trial++) {
        // ---- corner test on smooth-ish noise
        const int amp = 1 + (int)(rnd() % 255);
        for (int i = 0; i < H * W; i++) img[i] = (uint8_t)(128 + (int)(rnd() % (unsigned)amp) - amp / 2);
        if (trial % 7 == 0) for (int i = 0; i < H * W; i++) img[i] = (rnd() & 1) ? 255 : 0;
        to_planes(img, planes);
        for (int r = 3; r < H - 3; r++)
            for (int s = 0; s < NSEG; s++) {
                const uint32_t got = bs_corners<PLP>(planes + r * PLP + (s + 1) * 8);
                for (int j = 0; j < 32; j++) {
                    const int x = 32 * s + j;
                    if (x < 3 || x >= W - 3) continue;
                    const int c = img[r * W + x];
                    int b = 0, d = 0;
                    for (int k = 0; k < 16; k++) {
                        const int v = img[(r + dy[k]) * W + x + dx[k]];
                        if (v > c + 20) b |= 1 << k;
                        if (v < c - 20) d |= 1 << k;
                    }
                    bool is = false;
                    for (int k = 0; k < 16 && !is; k++) {
                        const int m = ((0x1FF << k) | (0x1FF >> (16 - k))) & 0xFFFF;
                        is = (b & m) == m || (d & m) == m;
                    }
                    corners += is;
                    if (is != (((got >> j) & 1u) != 0)) bad++;
                }
            }
    }
    // ---- the paired formulation (bs_pairs + bs_corners_paired) must give the same corner words as bs_corners
    {
        const int GP = 16 * NSEG + 2;
        static uint32_t gbuf[16 + H * GP];
        uint32_t *g = gbuf + 16;
        for (int trial = 0; trial < 100; trial++) {
            const int amp = 1 + (int)(rnd() % 255);
            for (int i = 0; i < H * W; i++) img[i] = (uint8_t)(128 + (int)(rnd() % (unsigned)amp) - amp / 2);
            if (trial % 5 == 0) for (int i = 0; i < H * W; i++) img[i] = (rnd() & 1) ? 255 : 0;
            to_planes(img, planes);
            static uint32_t Gp[H][NSEG][8], Gm[H][NSEG][8];
            memset(gbuf, 0xA5, sizeof(gbuf));          // pair words of rows without a full ring are never read for a valid pixel
            for (int r = 3; r < H - 3; r++)            // (the kernel pads its plane buffer by three rows instead)
                for (int s2 = 0; s2 < NSEG; s2++) {
                    bs_pairs<PLP>(planes + r * PLP + (s2 + 1) * 8, Gp[r][s2], Gm[r][s2]);
                    for (int k = 0; k < 8; k++) { g[r * GP + s2 * 16 + 2 * k] = Gp[r][s2][k]; g[r * GP + s2 * 16 + 2 * k + 1] = Gm[r][s2][k]; }
                }
            for (int r = 6; r < H - 6; r++)
                for (int s2 = 0; s2 < NSEG; s2++) {
                    const uint32_t want = bs_corners<PLP>(planes + r * PLP + (s2 + 1) * 8);
                    const uint32_t got = bs_corners_paired<GP>(Gp[r][s2], Gm[r][s2], g + r * GP + s2 * 16);
                    uint32_t valid = 0xFFFFFFFFu;
                    if (s2 == 0) valid &= ~7u;                      // columns 0..2: no full ring
                    if (s2 == NSEG - 1) valid &= 0x1FFFFFFFu;       // columns W-3..W-1
                    if ((got ^ want) & valid) bad++;
                }
        }
    }
    printf("corners %ld mismatches %ld\n", corners, bad);
    return bad ? 1 : 0;
}
