/* A plain C host for libbbx.so: no Python, no torch -- cudaMalloc'd buffers, the C ABI of
 * include/bbx.h, and a scalar check written out longhand.  Built and run by
 * tests/test_cabi_host_gpu.py (gcc host_c_abi.c -I include -lbbx -lcudart).
 *
 *   1. bbx_stack_median   (master_prep core, blackbox.py:4908-4984 + 5063-5073): N = 7 flats with
 *      divisors, the edge / non-positive post-fix, against a sorted-array median per pixel;
 *   2. bbx_xtalk          (xtalk_corr, blackbox.py:7138-7258) on a 2 x 8 channel frame against
 *      the per-pixel sum over the 15 source channels (float64 accumulation, mirrored rows for the
 *      channels of the other read-out half);
 *   3. bbx_fits_decode / bbx_fits_encode round trip of unsigned 16-bit counts;
 *   4. an argument error returns < 0 and bbx_last_error() names the entry point.
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bbx.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "%s:%d CUDA %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 2; } } while (0)
#define CHECK_BBX(x) do { if ((x) != 0) { \
    fprintf(stderr, "%s:%d bbx: %s\n", __FILE__, __LINE__, bbx_last_error()); return 3; } } while (0)

static uint32_t lcg_state = 12345u;
static float unit(void) { lcg_state = lcg_state * 1664525u + 1013904223u; return (float)(lcg_state >> 8) / 16777216.0f; }

static int cmp_float(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

static int test_stack_median(void) {
    enum { N = 7, H = 64, W = 520 };
    const size_t npix = (size_t)H * W;
    const int edge = 32;
    float *host[N], *dev[N], scale[N], *out_d, *out_h = malloc(npix * 4);
    uint8_t *bpm_h = calloc(npix, 1), *bpm_d;
    for (int i = 0; i < N; i++) {
        host[i] = malloc(npix * 4);
        scale[i] = (i == 3) ? 0.0f : 0.8f + 0.1f * i;           /* 0 = leave unscaled */
        for (size_t p = 0; p < npix; p++) host[i][p] = 0.5f + unit();
        host[i][17] = -1.0f;                                    /* median <= 0 -> 1 */
        CHECK_CUDA(cudaMalloc((void **)&dev[i], npix * 4));
        CHECK_CUDA(cudaMemcpy(dev[i], host[i], npix * 4, cudaMemcpyHostToDevice));
    }
    for (int x = 0; x < W; x++) bpm_h[x] = (uint8_t)edge;      /* first row: edge -> 1 */
    bpm_h[5 * W + 9] = 1;                                       /* a bad pixel is left alone */
    CHECK_CUDA(cudaMalloc((void **)&bpm_d, npix));
    CHECK_CUDA(cudaMemcpy(bpm_d, bpm_h, npix, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMalloc((void **)&out_d, npix * 4));
    CHECK_BBX(bbx_stack_median((const float *const *)dev, scale, N, npix, 1, bpm_d, edge, out_d, NULL));
    CHECK_CUDA(cudaDeviceSynchronize());
    CHECK_CUDA(cudaMemcpy(out_h, out_d, npix * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t p = 0; p < npix; p++) {
        float v[N];
        for (int i = 0; i < N; i++) v[i] = scale[i] != 0.0f ? host[i][p] / scale[i] : host[i][p];
        qsort(v, N, sizeof(float), cmp_float);
        float want = v[N / 2];
        if (bpm_h[p] == edge || want <= 0.0f) want = 1.0f;
        if (memcmp(&want, &out_h[p], 4) != 0) bad++;
    }
    printf("stack_median: %zu of %zu pixels differ\n", bad, npix);
    for (int i = 0; i < N; i++) { cudaFree(dev[i]); free(host[i]); }
    cudaFree(bpm_d); cudaFree(out_d); free(bpm_h); free(out_h);
    return bad != 0;
}

static int test_xtalk(void) {
    enum { YS = 24, XS = 40, H = 2 * YS, W = 8 * XS };
    const bbx_maskbits bits = {1, 2, 4, 8, 16, 32, 64};
    static float img[H][W], got[H][W];
    static uint8_t mask[H][W];
    static double coef[16][16];                                 /* [source][victim] */
    for (int s = 0; s < 16; s++)
        for (int v = 0; v < 16; v++) coef[s][v] = (s == v) ? 0.0 : (unit() - 0.5) * 6e-4;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            img[y][x] = 2000.0f * unit() - 100.0f;              /* some pixels <= 0: not a source */
            float r = unit();
            mask[y][x] = r < 0.02f ? 1 : r < 0.04f ? 2 : r < 0.06f ? 32 : r < 0.08f ? 4 : 0;
        }
    float *img_d; uint8_t *mask_d;
    CHECK_CUDA(cudaMalloc((void **)&img_d, sizeof img));
    CHECK_CUDA(cudaMalloc((void **)&mask_d, sizeof mask));
    CHECK_CUDA(cudaMemcpy(img_d, img, sizeof img, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(mask_d, mask, sizeof mask, cudaMemcpyHostToDevice));
    CHECK_BBX(bbx_xtalk(img_d, mask_d, H, W, YS, XS, &coef[0][0], &bits, NULL));
    CHECK_CUDA(cudaDeviceSynchronize());
    CHECK_CUDA(cudaMemcpy(got, img_d, sizeof got, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    double worst = 0.0;
    for (int v = 0; v < 16; v++)
        for (int y = 0; y < YS; y++)
            for (int x = 0; x < XS; x++) {
                int vy = (v / 8) * YS + y, vx = (v % 8) * XS + x;
                double corr = 0.0;
                for (int s = 0; s < 16; s++) {
                    int ly = (s / 8 == v / 8) ? y : YS - 1 - y;  /* other half: rows mirrored */
                    int sy = (s / 8) * YS + ly, sx = (s % 8) * XS + x;
                    int ok = img[sy][sx] > 0.0f && !(mask[sy][sx] & (bits.bad | bits.cosmic));
                    if (ok) corr += coef[s][v] * (double)img[sy][sx];
                }
                float want = (mask[vy][vx] & bits.edge) ? img[vy][vx] : (float)((double)img[vy][vx] - corr);
                double d = fabs((double)want - (double)got[vy][vx]);
                /* the reference sums with np.matmul (order unspecified): float class, 1e-5 */
                if (d > 1e-5 * fabs((double)want) + 1e-5) bad++;
                if (d > worst) worst = d;
            }
    printf("xtalk: %zu pixels outside the float class, worst |diff| %.3g\n", bad, worst);
    cudaFree(img_d); cudaFree(mask_d);
    return bad != 0;
}

static int test_fits_codec(void) {
    enum { NPX = 4096 };
    uint8_t be[NPX * 2], back[NPX * 2], *be_d, *back_d;
    uint16_t counts[NPX], *cnt_d;
    for (int i = 0; i < NPX; i++) {
        uint16_t c = (uint16_t)(unit() * 65535.0f);
        int16_t stored = (int16_t)((int)c - 32768);             /* BZERO = 32768 */
        be[2 * i] = (uint8_t)((uint16_t)stored >> 8);
        be[2 * i + 1] = (uint8_t)((uint16_t)stored & 0xff);
    }
    CHECK_CUDA(cudaMalloc((void **)&be_d, sizeof be));
    CHECK_CUDA(cudaMalloc((void **)&back_d, sizeof back));
    CHECK_CUDA(cudaMalloc((void **)&cnt_d, sizeof counts));
    CHECK_CUDA(cudaMemcpy(be_d, be, sizeof be, cudaMemcpyHostToDevice));
    CHECK_BBX(bbx_fits_decode(be_d, 16, 1, NPX, cnt_d, NULL));
    CHECK_BBX(bbx_fits_encode(cnt_d, 16, 1, NPX, back_d, NULL));
    CHECK_CUDA(cudaDeviceSynchronize());
    CHECK_CUDA(cudaMemcpy(counts, cnt_d, sizeof counts, cudaMemcpyDeviceToHost));
    CHECK_CUDA(cudaMemcpy(back, back_d, sizeof back, cudaMemcpyDeviceToHost));
    size_t bad = memcmp(be, back, sizeof be) != 0;
    for (int i = 0; i < NPX; i++) {
        int stored = (int16_t)(uint16_t)((be[2 * i] << 8) | be[2 * i + 1]);
        if ((int)counts[i] != stored + 32768) bad++;
    }
    printf("fits codec: %zu mismatches\n", bad);
    cudaFree(be_d); cudaFree(back_d); cudaFree(cnt_d);
    return bad != 0;
}

static int test_errors(void) {
    int rc = bbx_stack_median(NULL, NULL, 0, 0, 0, NULL, 0, NULL, NULL);
    int ok = rc < 0 && strstr(bbx_last_error(), "bbx_stack_median") != NULL;
    printf("error path: rc %d, \"%s\"\n", rc, bbx_last_error());
    return !ok;
}

int main(void) {
    printf("libbbx version %d\n", bbx_version());
    int fails = 0, rc;
    if ((rc = test_errors()) != 0) { fails++; }
    if ((rc = test_stack_median()) != 0) { if (rc > 1) return rc; fails++; }
    if ((rc = test_xtalk()) != 0) { if (rc > 1) return rc; fails++; }
    if ((rc = test_fits_codec()) != 0) { if (rc > 1) return rc; fails++; }
    printf(fails ? "FAILED %d\n" : "C ABI host: all ok\n", fails);
    return fails != 0;
}
