// Host-only check of the in-register DFT templates against a naive double DFT (no GPU needed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../../sound-event-localization-detection_b200/csrc/dft_inreg.cuh"
using namespace seld;

template <int N, bool INV>
double check() {
    float2 v[N];
    double xr[N], xi[N];
    srand(N * 7 + INV);
    for (int i = 0; i < N; ++i) {
        xr[i] = rand() / double(RAND_MAX) - 0.5;
        xi[i] = rand() / double(RAND_MAX) - 0.5;
        v[i] = make_float2((float)xr[i], (float)xi[i]);
        xr[i] = v[i].x; xi[i] = v[i].y;
    }
    Dft<N, INV>::run(v);
    double err = 0;
    for (int k = 0; k < N; ++k) {
        double sr = 0, si = 0;
        for (int n = 0; n < N; ++n) {
            double a = (INV ? 2.0 : -2.0) * M_PI * double((long long)n * k % N) / N;
            sr += xr[n] * cos(a) - xi[n] * sin(a);
            si += xr[n] * sin(a) + xi[n] * cos(a);
        }
        err = fmax(err, fmax(fabs(sr - v[k].x), fabs(si - v[k].y)));
    }
    printf("N=%2d inv=%d max abs err %.3e\n", N, (int)INV, err);
    return err;
}

int main() {
    double e = 0;
    e = fmax(e, check<2, false>()); e = fmax(e, check<3, false>()); e = fmax(e, check<4, false>());
    e = fmax(e, check<5, false>()); e = fmax(e, check<6, false>()); e = fmax(e, check<8, false>());
    e = fmax(e, check<10, false>()); e = fmax(e, check<15, false>()); e = fmax(e, check<16, false>());
    e = fmax(e, check<30, false>()); e = fmax(e, check<32, false>());
    e = fmax(e, check<3, true>()); e = fmax(e, check<5, true>()); e = fmax(e, check<30, true>());
    e = fmax(e, check<32, true>()); e = fmax(e, check<16, true>());
    if (e > 5e-6) { printf("FAIL\n"); return 1; }
    printf("OK\n");
    return 0;
}
