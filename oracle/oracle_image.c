/* oracle_image.c -- CPU restatement of the image side channel's device stage.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * The reference's ring (src/netFPGA.cpp:292-365) pushes single-channel frames of original_h * original_w bytes through a
 * device task `image_process` (:303, :326) whose source and bitstream are absent: "parity unpinned" for its arithmetic.
 * Pinned by the reference: one byte per pixel in, one byte per pixel out, same size (:441-442, :321, :328), FIFO order (:322,
 * :355).  The filter itself is a builder decision: 3 x 3 binomial smoothing, replicated borders, integers:
 *     out[y][x] = (sum_{dy,dx in -1..1} w[dy] w[dx] in[clamp(y+dy)][clamp(x+dx)] + 8) >> 4,   w = (1, 2, 1). */
#include "oracle.h"

void oracle_filter3x3(const uint8_t *in, uint8_t *out, int h, int w)
{
    static const int wt[3] = {1, 2, 1};
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
        {
            int acc = 0;
            for (int dy = -1; dy <= 1; dy++)
            {
                int yy = y + dy;
                yy = yy < 0 ? 0 : (yy >= h ? h - 1 : yy);
                for (int dx = -1; dx <= 1; dx++)
                {
                    int xx = x + dx;
                    xx = xx < 0 ? 0 : (xx >= w ? w - 1 : xx);
                    acc += wt[dy + 1] * wt[dx + 1] * (int)in[(size_t)yy * (size_t)w + (size_t)xx];
                }
            }
            out[(size_t)y * (size_t)w + (size_t)x] = (uint8_t)((acc + 8) >> 4);
        }
}
