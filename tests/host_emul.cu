// CPU lane-by-lane emulation of the unit FFT data flow in csrc/fft_core.cuh.
// Built and run by tests/test_host_emul.py (nvcc host compilation, no GPU needed).
// Checks, for NF = 512 and 1024, against a float64 direct DFT:
//   (1) forward FFT of two packed real frames + split  == rfft of each frame
//   (2) merge + inverse FFT of two Hermitian spectra    == irfft of each (x NF)
//   (3) bin_of() covers every one-sided bin exactly once
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>
#include "../xai-audio-deepfakes_b200/csrc/fft_core.cuh"
#include "../xai-audio-deepfakes_b200/csrc/fft3.cuh"

using namespace adv;
typedef std::complex<double> cd;
struct Tw {
    const float2* t;
    __host__ __device__ float2 operator()(int k) const { return t[k]; }
};

template <int NF>
int run() {
    using G = Geo<NF>;
    const int L = G::LANES;
    std::vector<float> xa(NF), xb(NF);
    srand(7 + NF);
    for (int i = 0; i < NF; ++i) {
        xa[i] = (float)rand() / RAND_MAX - 0.5f;
        xb[i] = (float)rand() / RAND_MAX - 0.5f;
    }
    // float64 reference rfft
    std::vector<cd> RA(NF / 2 + 1), RB(NF / 2 + 1);
    for (int k = 0; k <= NF / 2; ++k) {
        cd sa = 0, sb = 0;
        for (int n = 0; n < NF; ++n) {
            cd w = std::polar(1.0, -2.0 * M_PI * (double)((long)k * n % NF) / NF);
            sa += (double)xa[n] * w;
            sb += (double)xb[n] * w;
        }
        RA[k] = sa;
        RB[k] = sb;
    }
    // per-lane twiddles
    std::vector<std::vector<float2>> tw(L, std::vector<float2>(32));
    for (int l = 0; l < L; ++l)
        for (int k1 = 0; k1 < 32; ++k1) {
            double a = -2.0 * M_PI * (double)(l * k1) / NF;
            tw[l][k1] = make_float2((float)cos(a), (float)sin(a));
        }
    std::vector<std::vector<float2>> V(L, std::vector<float2>(32));
    std::vector<float> scratch(G::SCRATCH);
    for (int l = 0; l < L; ++l)
        for (int n1 = 0; n1 < 32; ++n1) V[l][n1] = make_float2(xa[n1 * G::R2 + l], xb[n1 * G::R2 + l]);
    for (int l = 0; l < L; ++l) fwd_cols<NF>(V[l].data(), Tw{tw[l].data()});
    for (int im = 0; im < 2; ++im) {
        for (int l = 0; l < L; ++l) scr_store_cols<NF>(V[l].data(), l, scratch.data(), im);
        for (int l = 0; l < L; ++l) scr_load_rows<NF>(V[l].data(), l, scratch.data(), im);
    }
    for (int l = 0; l < L; ++l) fwd_rows<NF>(V[l].data());

    // split
    std::vector<std::vector<float2>> XA(L, std::vector<float2>(17)), XB(L, std::vector<float2>(17));
    if constexpr (NF == 512) {
        for (int l = 0; l < L; ++l) split512(V[l].data(), l, XA[l].data(), XB[l].data());
    } else {
        std::vector<std::vector<float2>> send(L, std::vector<float2>(16));
        for (int l = 0; l < L; ++l) split1024_pre(V[l].data(), send[l].data());
        for (int l = 0; l < L; ++l)
            split1024_post(V[l].data(), l, send[(32 - l) & 31].data(), XA[l].data(), XB[l].data());
    }
    double err = 0, mx = 0;
    std::vector<int> seen(NF / 2 + 1, 0);
    for (int l = 0; l < L; ++l)
        for (int i = 0; i < 17; ++i) {
            int b = bin_of<NF>(l, i);
            if (b < 0) continue;
            if (b > NF / 2) { printf("bin out of range %d\n", b); return 1; }
            seen[b]++;
            err = fmax(err, std::abs(cd(XA[l][i].x, XA[l][i].y) - RA[b]));
            err = fmax(err, std::abs(cd(XB[l][i].x, XB[l][i].y) - RB[b]));
            mx = fmax(mx, std::abs(RA[b]));
        }
    for (int k = 0; k <= NF / 2; ++k)
        if (seen[k] != 1) { printf("NF=%d bin %d seen %d times\n", NF, k, seen[k]); return 1; }
    printf("NF=%d forward+split max err %.3e (max |X| %.3f)\n", NF, err, mx);
    if (err > 2e-5 * mx) return 1;

    // merge + inverse: feed back XA, XB (with garbage in Im of DC/Nyquist to test C2R semantics)
    for (int l = 0; l < L; ++l)
        for (int i = 0; i < 17; ++i) {
            int b = bin_of<NF>(l, i);
            if (b == 0 || b == NF / 2) { XA[l][i].y = 123.0f; XB[l][i].y = -77.0f; }
            if (b < 0) { XA[l][i] = make_float2(0, 0); XB[l][i] = make_float2(0, 0); }
        }
    if constexpr (NF == 512) {
        for (int l = 0; l < L; ++l) merge512(V[l].data(), l, XA[l].data(), XB[l].data());
    } else {
        std::vector<std::vector<float2>> send(L, std::vector<float2>(16));
        for (int l = 0; l < L; ++l) merge1024_pre(V[l].data(), l, XA[l].data(), XB[l].data(), send[l].data());
        for (int l = 0; l < L; ++l) merge1024_post(V[l].data(), l, send[(32 - l) & 31].data());
    }
    for (int l = 0; l < L; ++l) inv_rows<NF>(V[l].data());
    for (int im = 0; im < 2; ++im) {
        for (int l = 0; l < L; ++l) scr_store_rows<NF>(V[l].data(), l, scratch.data(), im);
        for (int l = 0; l < L; ++l) scr_load_cols<NF>(V[l].data(), l, scratch.data(), im);
    }
    for (int l = 0; l < L; ++l) inv_cols<NF>(V[l].data(), Tw{tw[l].data()});
    double ierr = 0;
    for (int l = 0; l < L; ++l)
        for (int n1 = 0; n1 < 32; ++n1) {
            int n = n1 * G::R2 + l;
            ierr = fmax(ierr, fabs(V[l][n1].x / NF - xa[n]));
            ierr = fmax(ierr, fabs(V[l][n1].y / NF - xb[n]));
        }
    printf("NF=%d merge+inverse round-trip max err %.3e\n", NF, ierr);
    return ierr > 2e-6 ? 1 : 0;
}

// wide 512-point unit (32 lanes x 16 values): same checks
int run_w512() {
    const int NF = 512, L = 32;
    std::vector<float> xa(NF), xb(NF);
    srand(99);
    for (int i = 0; i < NF; ++i) {
        xa[i] = (float)rand() / RAND_MAX - 0.5f;
        xb[i] = (float)rand() / RAND_MAX - 0.5f;
    }
    std::vector<cd> RA(NF / 2 + 1), RB(NF / 2 + 1);
    for (int k = 0; k <= NF / 2; ++k) {
        cd sa = 0, sb = 0;
        for (int n = 0; n < NF; ++n) {
            cd w = std::polar(1.0, -2.0 * M_PI * (double)((long)k * n % NF) / NF);
            sa += (double)xa[n] * w;
            sb += (double)xb[n] * w;
        }
        RA[k] = sa;
        RB[k] = sb;
    }
    std::vector<std::vector<float2>> tw(L, std::vector<float2>(16)), V(L, std::vector<float2>(16)), O(L, std::vector<float2>(16));
    for (int l = 0; l < L; ++l)
        for (int k1 = 0; k1 < 16; ++k1) {
            double a = -2.0 * M_PI * (double)(l * k1) / NF;
            tw[l][k1] = make_float2((float)cos(a), (float)sin(a));
        }
    std::vector<float> scratch(w512::SCRATCH);
    for (int l = 0; l < L; ++l)
        for (int n1 = 0; n1 < 16; ++n1) V[l][n1] = make_float2(xa[n1 * 32 + l], xb[n1 * 32 + l]);
    for (int l = 0; l < L; ++l) w512::fwd_cols(V[l].data(), Tw{tw[l].data()});
    for (int im = 0; im < 2; ++im) {
        for (int l = 0; l < L; ++l) w512::scr_store_cols(V[l].data(), l, scratch.data(), im);
        for (int l = 0; l < L; ++l) w512::scr_load_rows(V[l].data(), l, scratch.data(), im);
    }
    for (int l = 0; l < L; ++l) w512::fwd_rows_local(V[l].data());
    O = V;
    for (int l = 0; l < L; ++l) w512::fwd_rows_combine(V[l].data(), l, O[l ^ 1].data());
    {   // lean pair butterfly (twiddle on the odd lane, then v = other + sg v) must agree with the reference form
        auto T = O, U = O;
        for (int l = 0; l < L; ++l) w512::fwd_rows_tw(T[l].data(), l);
        U = T;
        for (int l = 0; l < L; ++l) w512::rows_bfly(T[l].data(), l, U[l ^ 1].data());
        double d = 0;
        for (int l = 0; l < L; ++l)
            for (int m = 0; m < 16; ++m)
                d = fmax(d, hypot((double)T[l][m].x - V[l][m].x, (double)T[l][m].y - V[l][m].y));
        printf("W512 lean forward butterfly vs reference form: %.3e\n", d);
        if (d > 1e-4) return 1;
    }
    // split
    std::vector<std::vector<float2>> send(L, std::vector<float2>(8)), XA(L, std::vector<float2>(9)), XB(L, std::vector<float2>(9));
    for (int l = 0; l < L; ++l) w512::split_pre(V[l].data(), send[l].data());
    for (int l = 0; l < L; ++l) w512::split_post(V[l].data(), l, send[w512::partner_row(l)].data(), XA[l].data(), XB[l].data());
    double err = 0, mx = 0;
    std::vector<int> seen(NF / 2 + 1, 0);
    for (int l = 0; l < L; ++l)
        for (int i = 0; i < 9; ++i) {
            int b = w512::bin_of(l, i);
            if (b < 0) continue;
            if (b > NF / 2) { printf("w512 bin out of range %d (lane %d slot %d)\n", b, l, i); return 1; }
            seen[b]++;
            err = fmax(err, std::abs(cd(XA[l][i].x, XA[l][i].y) - RA[b]));
            err = fmax(err, std::abs(cd(XB[l][i].x, XB[l][i].y) - RB[b]));
            mx = fmax(mx, std::abs(RA[b]));
        }
    for (int k = 0; k <= NF / 2; ++k)
        if (seen[k] != 1) { printf("w512 bin %d seen %d times\n", k, seen[k]); return 1; }
    printf("W512 forward+split max err %.3e (max |X| %.3f)\n", err, mx);
    if (err > 2e-5 * mx) return 1;
    // merge + inverse
    for (int l = 0; l < L; ++l)
        for (int i = 0; i < 9; ++i) {
            int b = w512::bin_of(l, i);
            if (b == 0 || b == NF / 2) { XA[l][i].y = 123.0f; XB[l][i].y = -77.0f; }
            if (b < 0) { XA[l][i] = make_float2(0, 0); XB[l][i] = make_float2(0, 0); }
        }
    for (int l = 0; l < L; ++l) w512::merge_pre(V[l].data(), l, XA[l].data(), XB[l].data(), send[l].data());
    for (int l = 0; l < L; ++l) w512::merge_post(V[l].data(), l, send[w512::partner_row(l)].data());
    O = V;
    for (int l = 0; l < L; ++l) w512::inv_rows_combine(V[l].data(), l, O[l ^ 1].data());
    {
        auto T = O;
        for (int l = 0; l < L; ++l) w512::rows_bfly(T[l].data(), l, O[l ^ 1].data());
        for (int l = 0; l < L; ++l) w512::inv_rows_tw(T[l].data(), l);
        double d = 0;
        for (int l = 0; l < L; ++l)
            for (int m = 0; m < 16; ++m)
                d = fmax(d, hypot((double)T[l][m].x - V[l][m].x, (double)T[l][m].y - V[l][m].y));
        printf("W512 lean inverse butterfly vs reference form: %.3e\n", d);
        if (d > 1e-3) return 1;
    }
    for (int l = 0; l < L; ++l) w512::inv_rows_local(V[l].data());
    for (int im = 0; im < 2; ++im) {
        for (int l = 0; l < L; ++l) w512::scr_store_rows(V[l].data(), l, scratch.data(), im);
        for (int l = 0; l < L; ++l) w512::scr_load_cols(V[l].data(), l, scratch.data(), im);
    }
    for (int l = 0; l < L; ++l) w512::inv_cols(V[l].data(), Tw{tw[l].data()});
    double ierr = 0;
    for (int l = 0; l < L; ++l)
        for (int n1 = 0; n1 < 16; ++n1) {
            int n = n1 * 32 + l;
            ierr = fmax(ierr, fabs(V[l][n1].x / NF - xa[n]));
            ierr = fmax(ierr, fabs(V[l][n1].y / NF - xb[n]));
        }
    printf("W512 merge+inverse round-trip max err %.3e\n", ierr);
    return ierr > 2e-6 ? 1 : 0;
}

// generation-3 core (fft3.cuh): 8 x 8 x 8, shuffle-free, lane-local mirror bins.  VEC: 64-bit / planar exchanges.
template <bool VEC>
void f3_forward(std::vector<std::vector<float2>>& V, const float2* tw, std::vector<float>& scr) {
    using E = f3::X<VEC>;
    const int L = 32, planes = VEC ? 1 : 2;
    for (int l = 0; l < L; ++l) f3::fwd_a(V[l].data(), l, tw);
    for (int im = 0; im < planes; ++im) {
        for (int l = 0; l < L; ++l) E::x1_store_cols(V[l].data(), l, scr.data(), im);
        for (int l = 0; l < L; ++l) E::x1_load_rows(V[l].data(), l, scr.data(), im);
    }
    for (int l = 0; l < L; ++l) f3::fwd_b(V[l].data(), l, tw);
    for (int im = 0; im < planes; ++im) {
        for (int l = 0; l < L; ++l) E::x2_store_rows(V[l].data(), l, scr.data(), im);
        for (int l = 0; l < L; ++l) E::x2_load_cols(V[l].data(), l, scr.data(), im);
    }
    for (int l = 0; l < L; ++l) f3::fwd_c(V[l].data());
}
template <bool VEC>
void f3_inverse(std::vector<std::vector<float2>>& V, const float2* tw, std::vector<float>& scr) {
    using E = f3::X<VEC>;
    const int L = 32, planes = VEC ? 1 : 2;
    for (int l = 0; l < L; ++l) f3::inv_c(V[l].data(), l, tw);
    for (int im = 0; im < planes; ++im) {
        for (int l = 0; l < L; ++l) E::x2_store_cols(V[l].data(), l, scr.data(), im);
        for (int l = 0; l < L; ++l) E::x2_load_rows(V[l].data(), l, scr.data(), im);
    }
    for (int l = 0; l < L; ++l) f3::inv_b(V[l].data(), l, tw);
    for (int im = 0; im < planes; ++im) {
        for (int l = 0; l < L; ++l) E::x1_store_rows(V[l].data(), l, scr.data(), im);
        for (int l = 0; l < L; ++l) E::x1_load_cols(V[l].data(), l, scr.data(), im);
    }
    for (int l = 0; l < L; ++l) f3::inv_a(V[l].data());
}

template <bool VEC>
int run_f3() {
    const int NF = 512, L = 32;
    std::vector<float2> tw(f3::TW_TOTAL);
    f3::build_tables(tw.data());
    std::vector<float> scr(f3::Scr<VEC>::FLOATS + 64, 0.f);
    // ---- (1) two packed real frames: forward + split == rfft of each; merge + inverse round trip
    std::vector<float> xa(NF), xb(NF);
    srand(1234 + VEC);
    for (int i = 0; i < NF; ++i) {
        xa[i] = (float)rand() / RAND_MAX - 0.5f;
        xb[i] = (float)rand() / RAND_MAX - 0.5f;
    }
    std::vector<cd> RA(NF / 2 + 1), RB(NF / 2 + 1);
    for (int k = 0; k <= NF / 2; ++k) {
        cd sa = 0, sb = 0;
        for (int n = 0; n < NF; ++n) {
            cd w = std::polar(1.0, -2.0 * M_PI * (double)((long)k * n % NF) / NF);
            sa += (double)xa[n] * w;
            sb += (double)xb[n] * w;
        }
        RA[k] = sa;
        RB[k] = sb;
    }
    std::vector<std::vector<float2>> V(L, std::vector<float2>(16));
    for (int l = 0; l < L; ++l)
        for (int j = 0; j < 16; ++j) V[l][j] = make_float2(xa[32 * j + l], xb[32 * j + l]);
    f3_forward<VEC>(V, tw.data(), scr);
    // raw spectrum placement: register 8 g + k0 of lane l is Z[q_g + 64 k0]
    {
        double e = 0;
        for (int l = 0; l < L; ++l)
            for (int g = 0; g < 2; ++g)
                for (int k0 = 0; k0 < 8; ++k0) {
                    const int k = (g ? f3::q1_of(l) : f3::q0_of(l)) + 64 * k0;
                    cd z = 0;
                    for (int n = 0; n < NF; ++n)
                        z += cd(xa[n], xb[n]) * std::polar(1.0, -2.0 * M_PI * (double)((long)k * n % NF) / NF);
                    e = fmax(e, std::abs(cd(V[l][8 * g + k0].x, V[l][8 * g + k0].y) - z));
                }
        printf("F3<%d> forward placement max err %.3e\n", (int)VEC, e);
        if (e > 3e-4) return 1;
    }
    std::vector<std::vector<float2>> XA(L, std::vector<float2>(9)), XB(L, std::vector<float2>(9));
    for (int l = 0; l < L; ++l) f3::split(V[l].data(), l, XA[l].data(), XB[l].data());
    double err = 0, mx = 0;
    std::vector<int> seen(NF / 2 + 1, 0);
    for (int l = 0; l < L; ++l)
        for (int i = 0; i < 9; ++i) {
            int b = f3::bin_of(l, i);
            if (b < 0) continue;
            if (b > NF / 2) { printf("f3 bin out of range %d (lane %d slot %d)\n", b, l, i); return 1; }
            seen[b]++;
            err = fmax(err, std::abs(cd(XA[l][i].x, XA[l][i].y) - RA[b]));
            err = fmax(err, std::abs(cd(XB[l][i].x, XB[l][i].y) - RB[b]));
            mx = fmax(mx, std::abs(RA[b]));
        }
    for (int k = 0; k <= NF / 2; ++k)
        if (seen[k] != 1) { printf("f3 bin %d seen %d times\n", k, seen[k]); return 1; }
    printf("F3<%d> forward+split max err %.3e (max |X| %.3f)\n", (int)VEC, err, mx);
    if (err > 2e-5 * mx) return 1;
    for (int l = 0; l < L; ++l)
        for (int i = 0; i < 9; ++i) {
            int b = f3::bin_of(l, i);
            if (b == 0 || b == NF / 2) { XA[l][i].y = 123.0f; XB[l][i].y = -77.0f; }   // C2R: imaginary parts ignored
            if (b < 0) { XA[l][i] = make_float2(9, 9); XB[l][i] = make_float2(-9, 9); } // empty slots must not matter
        }
    for (int l = 0; l < L; ++l) f3::merge(V[l].data(), l, XA[l].data(), XB[l].data());
    f3_inverse<VEC>(V, tw.data(), scr);
    double ierr = 0;
    for (int l = 0; l < L; ++l)
        for (int j = 0; j < 16; ++j) {
            ierr = fmax(ierr, fabs(V[l][j].x / NF - xa[32 * j + l]));
            ierr = fmax(ierr, fabs(V[l][j].y / NF - xb[32 * j + l]));
        }
    printf("F3<%d> merge+inverse round-trip max err %.3e\n", (int)VEC, ierr);
    if (ierr > 2e-6) return 1;

    // ---- (2) one real 1024-point frame through the 512-point transform
    const int N2 = 1024;
    std::vector<float> x(N2);
    for (int i = 0; i < N2; ++i) x[i] = (float)rand() / RAND_MAX - 0.5f;
    std::vector<cd> R(N2 / 2 + 1);
    for (int k = 0; k <= N2 / 2; ++k) {
        cd s = 0;
        for (int n = 0; n < N2; ++n) s += (double)x[n] * std::polar(1.0, -2.0 * M_PI * (double)((long)k * n % N2) / N2);
        R[k] = s;
    }
    for (int l = 0; l < L; ++l)
        for (int j = 0; j < 16; ++j) V[l][j] = make_float2(x[2 * (32 * j + l)], x[2 * (32 * j + l) + 1]);
    f3_forward<VEC>(V, tw.data(), scr);
    std::vector<std::vector<float2>> XK(L, std::vector<float2>(9)), XM(L, std::vector<float2>(9));
    for (int l = 0; l < L; ++l) f3::r1024_post(V[l].data(), l, tw.data(), XK[l].data(), XM[l].data());
    std::vector<int> seen2(N2 / 2 + 1, 0);
    double e2 = 0, m2 = 0;
    for (int l = 0; l < L; ++l)
        for (int i = 0; i < 9; ++i) {
            int k = f3::bin_of(l, i);
            if (k < 0) continue;
            seen2[k]++;
            e2 = fmax(e2, std::abs(cd(XK[l][i].x, XK[l][i].y) - R[k]));
            m2 = fmax(m2, std::abs(R[k]));
            if (k != 256) {
                seen2[512 - k]++;
                e2 = fmax(e2, std::abs(cd(XM[l][i].x, XM[l][i].y) - R[512 - k]));
            }
        }
    for (int k = 0; k <= N2 / 2; ++k)
        if (seen2[k] != 1) { printf("f3 r1024 bin %d seen %d times\n", k, seen2[k]); return 1; }
    printf("F3<%d> real-1024 forward max err %.3e (max |X| %.3f)\n", (int)VEC, e2, m2);
    if (e2 > 2e-5 * m2) return 1;
    for (int l = 0; l < L; ++l)
        for (int i = 0; i < 9; ++i) {
            int k = f3::bin_of(l, i);
            if (k == 0) { XK[l][i].y = 55.0f; XM[l][i].y = -31.0f; }      // imaginary parts of bins 0 and 512 ignored
            if (k == 256) XM[l][i] = XK[l][i];                            // slot 8 carries X[256] twice
            if (k < 0) { XK[l][i] = make_float2(7, -7); XM[l][i] = make_float2(3, 3); }
        }
    for (int l = 0; l < L; ++l) f3::r1024_pre(V[l].data(), l, tw.data(), XK[l].data(), XM[l].data());
    f3_inverse<VEC>(V, tw.data(), scr);
    double e3 = 0;
    for (int l = 0; l < L; ++l)
        for (int j = 0; j < 16; ++j) {
            e3 = fmax(e3, fabs(V[l][j].x / N2 - x[2 * (32 * j + l)]));
            e3 = fmax(e3, fabs(V[l][j].y / N2 - x[2 * (32 * j + l) + 1]));
        }
    printf("F3<%d> real-1024 inverse round-trip max err %.3e\n", (int)VEC, e3);
    return e3 > 2e-6 ? 1 : 0;
}

int main() {
    int rc = run<512>();
    rc |= run<1024>();
    rc |= run_w512();
    rc |= run_f3<true>();
    rc |= run_f3<false>();
    printf(rc ? "FAIL\n" : "OK\n");
    return rc;
}
