/* ORACLE — TEST INFRASTRUCTURE ONLY (oracle/README.md).  Plain-C scalar restatement of
 * StackedHistogram.construct (/root/reference/data/utils/representations.py:76-121), used
 * as a second, independent checker of oracle/voxel_oracle.py and as the single-thread CPU
 * baseline ("port") timed by bench.py.  Built by oracle/Makefile into oracle/_build/.
 *
 * The fp32 time-bin arithmetic mirrors torch's int64/int64 true-divide: convert both
 * operands to float (round-to-nearest-even), IEEE divide, multiply by bins, floor, clamp
 * (representations.py:102-109).  Compile WITHOUT -ffast-math.
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>

/* out: [2*bins*H*W] uint8.  Returns 0, or -1 on a bad argument (pol outside {0,1},
 * coordinates outside the frame, unsorted first/last timestamp). */
int rvt_oracle_stacked_histogram(const int64_t *x, const int64_t *y, const int64_t *pol,
                                 const int64_t *t, int64_t n, int bins, int height, int width,
                                 int count_cutoff, int fastmode, uint8_t *out)
{
    const int64_t hw = (int64_t)height * width;
    const int64_t n_out = 2 * (int64_t)bins * hw;
    int cutoff = count_cutoff <= 0 ? 255 : (count_cutoff > 255 ? 255 : count_cutoff);
    memset(out, 0, (size_t)n_out);
    if (n == 0) return 0;
    if (t[n - 1] < t[0]) return -1;
    int64_t dt = t[n - 1] - t[0];
    volatile float denom = (float)(dt > 1 ? dt : 1);
    if (fastmode) {
        for (int64_t i = 0; i < n; ++i) {
            if (pol[i] < 0 || pol[i] > 1 || x[i] < 0 || x[i] >= width || y[i] < 0 || y[i] >= height) return -1;
            volatile float q = (float)(t[i] - t[0]) / denom;
            volatile float s = q * (float)bins;
            float f = floorf(s);
            if (f > (float)(bins - 1)) f = (float)(bins - 1);
            int64_t idx = x[i] + width * y[i] + hw * (int64_t)f + bins * hw * pol[i];
            out[idx] = (uint8_t)(out[idx] + 1);          /* wraps mod 256 */
        }
        for (int64_t j = 0; j < n_out; ++j) if (out[j] > cutoff) out[j] = (uint8_t)cutoff;
    } else {
        int16_t *acc = (int16_t *)calloc((size_t)n_out, sizeof(int16_t));
        if (!acc) return -2;
        for (int64_t i = 0; i < n; ++i) {
            if (pol[i] < 0 || pol[i] > 1 || x[i] < 0 || x[i] >= width || y[i] < 0 || y[i] >= height) { free(acc); return -1; }
            volatile float q = (float)(t[i] - t[0]) / denom;
            volatile float s = q * (float)bins;
            float f = floorf(s);
            if (f > (float)(bins - 1)) f = (float)(bins - 1);
            int64_t idx = x[i] + width * y[i] + hw * (int64_t)f + bins * hw * pol[i];
            acc[idx] = (int16_t)(uint16_t)((uint16_t)acc[idx] + 1u);
        }
        for (int64_t j = 0; j < n_out; ++j) {
            int v = acc[j]; if (v < 0) v = 0; if (v > cutoff) v = cutoff;
            out[j] = (uint8_t)v;
        }
        free(acc);
    }
    return 0;
}
