/* Plain-C use of the C ABI (include/aig.h): no Python, no torch.  Builds the tables the way the reference does
 * (dataloader/outdoor_data_mfcc.py:806-849), runs MFCC -> energy -> IoU sweep on host buffers and prints a few values.
 *
 *   gcc -O2 -Iinclude examples/c_abi_smoke.c -o /tmp/c_abi_smoke -Lacoustic_image_generation_b200/csrc -laig -lm \
 *       -Wl,-rpath,$PWD/acoustic_image_generation_b200/csrc && /tmp/c_abi_smoke
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "aig.h"

#define FFT_LEN 512
#define FILTERS 24
#define MFCCS 12

static void linspace(double a, double b, int n, double* out) {
    /* numpy.linspace: a + i * step with the last point set exactly */
    const double step = n > 1 ? (b - a) / (n - 1) : 0.0;
    for (int i = 0; i < n; ++i) out[i] = a + i * step;
    if (n > 1) out[n - 1] = b;
}

int main(void) {
    static double bank[FFT_LEN * FILTERS], dct[FILTERS * MFCCS], lifter[MFCCS];
    double mel[FILTERS + 2], ramp[FFT_LEN];
    int edge[FILTERS + 2];
    const double pi = 3.14159265358979323846;
    linspace(1127.0 * log(1.0 + 0.0 / 700.0), 1127.0 * log(1.0 + 6400.0 / 700.0), FILTERS + 2, mel);
    for (int i = 0; i < FILTERS + 2; ++i)
        edge[i] = (int)floor(700.0 * (exp(mel[i] / 1127.0) - 1.0) / 12800.0 * (FFT_LEN - 1) * 2);
    for (int f = 0; f < FILTERS; ++f) {
        const int d1 = edge[f + 1] - edge[f], d2 = edge[f + 2] - edge[f + 1];
        linspace(0.0, 1.0, d1 + 1, ramp);
        for (int i = 0; i <= d1; ++i) bank[(edge[f] + i) * FILTERS + f] = ramp[i];
        linspace(1.0, 0.0, d2 + 1, ramp);
        for (int i = 0; i <= d2; ++i) bank[(edge[f + 1] + i) * FILTERS + f] = ramp[i];
    }
    for (int m = 0; m < MFCCS; ++m) {
        for (int j = 0; j < FILTERS; ++j) dct[j * MFCCS + m] = cos((m + 1) * pi / FILTERS * (j + 0.5));
        lifter[m] = 1 + (22 / 2.0) * sin(pi * (1 + m) / 22);
    }
    aig_handle* h = NULL;
    if (aig_create(0, 0, &h) != AIG_OK) { fprintf(stderr, "%s\n", aig_last_error(NULL)); return 1; }
    if (aig_set_tables(h, bank, FFT_LEN, FILTERS, dct, MFCCS, lifter, sqrt(2.0 / FILTERS)) != AIG_OK) {
        fprintf(stderr, "%s\n", aig_last_error(h)); return 1;
    }
    printf("tables match the fused kernel's reference configuration: %d\n", aig_tables_are_reference(h));
    const int frames = 2;
    const size_t rows = (size_t)frames * AIG_FRAME_PIXELS;
    float* power = malloc(rows * FFT_LEN * sizeof(float));
    float* mfcc = malloc(rows * MFCCS * sizeof(float));
    double* energy = malloc(rows * sizeof(double));
    unsigned char* mask = malloc(rows);
    unsigned s = 12345u;
    for (size_t i = 0; i < rows * FFT_LEN; ++i) { s = s * 1664525u + 1013904223u; float u = (s >> 8) * (1.0f / 16777216.0f); power[i] = u * u * 4.0f; }
    int rc;
    if (aig_tables_are_reference(h) == 1) {
        rc = aig_mfcc_energy(h, power, frames, 1, 1, mfcc, energy, mask, NULL);           /* fused kernel */
    } else {                              /* libm rounded a table entry differently from NumPy: generic float64 kernel */
        rc = aig_mfcc(h, power, (int64_t)rows, mfcc, 1, AIG_FRAME_PIXELS);
        if (rc == AIG_OK) rc = aig_energy(h, mfcc, frames, 1, NULL, energy, mask, NULL);
    }
    if (rc != AIG_OK) { fprintf(stderr, "%s\n", aig_last_error(h)); return 1; }
    double thr[11];
    int64_t pos[11] = {0}, num = 0, inter = 0, uni = 0;
    for (int k = 0; k < 11; ++k) thr[k] = k / 10.0;
    if (aig_iou_sweep(h, mask, mask + AIG_FRAME_PIXELS, 1, thr, 11, &inter, &uni, pos, &num) != AIG_OK) { fprintf(stderr, "%s\n", aig_last_error(h)); return 1; }
    double rate[11], area = 0.0;
    for (int k = 0; k < 11; ++k) rate[k] = (double)pos[k] / (double)num;
    aig_auc(thr, rate, 11, &area);
    printf("mfcc[0][0..2] = %.5f %.5f %.5f   energy[0] = %.9g   I = %lld U = %lld   auc = %.3f   launches = %lld\n", mfcc[0], mfcc[1], mfcc[2],
           energy[0], (long long)inter, (long long)uni, area, (long long)aig_launch_count(h));
    aig_destroy(h);
    free(power); free(mfcc); free(energy); free(mask);
    return 0;
}
