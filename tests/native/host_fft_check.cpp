// host_fft_check.cpp -- CPU check of the FFT correlation arithmetic (hdn_b200/csrc/xcorr_fft.cuh, fft64.cuh).
// Runs the kernel's three phases (row FFTs, column correlation, output FFTs) task by task on one group of planes -- in the order the kernel's barriers impose, once with the
// tasks of a phase in ascending and once in descending order (a result that depended on the order inside a phase would be a
// race on the device) -- and compares with a direct double-precision correlation.  Test infrastructure only: built and run
// by tests/test_fft_host.py with g++; prints "<config> max_err <e> max_ref <m>" per shape, exit code 1 on failure.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../hdn_b200/csrc/xcorr_fft.cuh"

using namespace hdn;

struct Regs {
    float re[32], im[32];
};

template <class Cfg>
static void run_group(const std::vector<float> &raw, std::vector<float> &out, bool reverse) {
    std::vector<float2> XR(Cfg::G * Cfg::XR_PLANE), KR(Cfg::G * Cfg::KR_PLANE), CT(Cfg::G * Cfg::CT_PLANE);
    for (auto &v : XR) v = float2{NAN, NAN};  // anything read before it is written poisons the result
    for (auto &v : KR) v = float2{NAN, NAN};
    for (auto &v : CT) v = float2{NAN, NAN};
    // the kernel stages the output tile over XR; here it is a separate array so that a phase-O read of XR would show up as NaN
    FftBufs b{raw.data(), raw.data() + Cfg::G * Cfg::XPL, XR.data(), KR.data(), CT.data(), out.data()};
    auto order = [&](int s, int n) { return reverse ? n - 1 - s : s; };
    auto fft_phase = [&](int ph) {
        const int ntask = fftc_tasks<Cfg>(ph);
        for (int s = 0; s < ntask; ++s) {
            const int t = order(s, ntask), h = fft_task_half<Cfg>(ph, t), unit = fft_task_unit<Cfg>(ph, t);
            Regs r;
            if (!fftc_load<Cfg>(ph, b, unit, h, r.re, r.im)) continue;
            if (h) fft::half_twiddle(r.re, r.im);
            fft::fft32_fwd(r.re, r.im);
            fftc_store<Cfg>(ph, b, unit, h, r.re, r.im);
        }
    };
    fft_phase(FFT_PH_R);
    for (int s = 0; s < Cfg::COL_TASKS; ++s) fftc_col<Cfg>(b, order(s, Cfg::COL_TASKS));
    fft_phase(FFT_PH_O);
}

template <class Cfg>
static int check(const char *name, unsigned seed) {
    std::vector<float> raw(Cfg::RAW_FLOATS), out(Cfg::OUT_FLOATS, -1.f), out2(Cfg::OUT_FLOATS, -2.f);
    srand(seed);
    for (auto &v : raw) v = (float)rand() / RAND_MAX * 2.f - 0.7f;  // non-zero mean, like post-ReLU features
    run_group<Cfg>(raw, out, false);
    run_group<Cfg>(raw, out2, true);
    const bool order_free = std::memcmp(out.data(), out2.data(), out.size() * sizeof(float)) == 0;
    const float *rawx = raw.data(), *rawk = raw.data() + Cfg::G * Cfg::XPL;
    double max_err = 0, max_ref = 0;
    for (int p = 0; p < Cfg::G; ++p)
        for (int i = 0; i < Cfg::HO; ++i)
            for (int j = 0; j < Cfg::WO; ++j) {
                double acc = 0;
                for (int u = 0; u < Cfg::KH; ++u) {
                    int r = i + u - Cfg::PH;
                    if (Cfg::CIRC) r = ((r % Cfg::HX) + Cfg::HX) % Cfg::HX;
                    for (int v = 0; v < Cfg::KW; ++v) {
                        int c = j + v - Cfg::PW;
                        c = c < 0 ? 0 : (c > Cfg::WX - 1 ? Cfg::WX - 1 : c);
                        acc += (double)rawx[p * Cfg::XPL + r * Cfg::WX + c] * rawk[p * Cfg::KPL + u * Cfg::KW + v];
                    }
                }
                const double e = std::fabs(acc - out[p * Cfg::OPL + i * Cfg::WO + j]);
                if (!(e <= max_err)) max_err = e;  // NaN-propagating
                if (std::fabs(acc) > max_ref) max_ref = std::fabs(acc);
            }
    printf("%s max_err %.3e max_ref %.3e order_free %d\n", name, max_err, max_ref, (int)order_free);
    return (max_err <= 2e-5 * max_ref && order_free) ? 0 : 1;
}

// Shared template with cached row spectra: the template's spectra are taken ONCE by the plain configuration's phase R with an all-zero
// x (what xcorr_spectra_kernel does), then the KSPEC configuration transforms x rows only and reads KR ready-made.
template <class Cfg, class SCfg>
static int check_spectra(const char *name, unsigned seed) {
    static_assert(!Cfg::KSPEC && SCfg::KSPEC && Cfg::KR_PLANE == SCfg::KR_PLANE && Cfg::XPL == SCfg::XPL, "matching configurations");
    std::vector<float> x(Cfg::G * Cfg::XPL), k(Cfg::G * Cfg::KPL), out(Cfg::OUT_FLOATS, -1.f);
    srand(seed);
    for (auto &v : x) v = (float)rand() / RAND_MAX * 2.f - 0.7f;
    for (auto &v : k) v = (float)rand() / RAND_MAX * 2.f - 0.7f;
    // producer: phase R of the plain configuration on (0, k)
    std::vector<float> raw0(Cfg::G * Cfg::XPL + Cfg::G * Cfg::KPL, 0.f);
    std::memcpy(raw0.data() + Cfg::G * Cfg::XPL, k.data(), k.size() * sizeof(float));
    std::vector<float2> XR0(Cfg::G * Cfg::XR_PLANE), KR(Cfg::G * Cfg::KR_PLANE, float2{NAN, NAN}), CT0(Cfg::G * Cfg::CT_PLANE);
    {
        FftBufs b{raw0.data(), raw0.data() + Cfg::G * Cfg::XPL, XR0.data(), KR.data(), CT0.data(), out.data()};
        for (int t = 0; t < Cfg::R_TASKS; ++t) {
            const int h = fft_task_half<Cfg>(FFT_PH_R, t), unit = fft_task_unit<Cfg>(FFT_PH_R, t);
            Regs r;
            if (!fftc_load<Cfg>(FFT_PH_R, b, unit, h, r.re, r.im)) continue;
            if (h) fft::half_twiddle(r.re, r.im);
            fft::fft32_fwd(r.re, r.im);
            fftc_store<Cfg>(FFT_PH_R, b, unit, h, r.re, r.im);
        }
    }
    // consumer: the KSPEC configuration's three phases with KR given
    std::vector<float2> XR(SCfg::G * SCfg::XR_PLANE, float2{NAN, NAN}), CT(SCfg::G * SCfg::CT_PLANE, float2{NAN, NAN});
    {
        FftBufs b{x.data(), nullptr, XR.data(), KR.data(), CT.data(), out.data()};
        auto fft_phase = [&](int ph) {
            const int ntask = fftc_tasks<SCfg>(ph);
            for (int t = 0; t < ntask; ++t) {
                const int h = fft_task_half<SCfg>(ph, t), unit = fft_task_unit<SCfg>(ph, t);
                Regs r;
                if (!fftc_load<SCfg>(ph, b, unit, h, r.re, r.im)) continue;
                if (h) fft::half_twiddle(r.re, r.im);
                fft::fft32_fwd(r.re, r.im);
                fftc_store<SCfg>(ph, b, unit, h, r.re, r.im);
            }
        };
        fft_phase(FFT_PH_R);
        for (int t = 0; t < SCfg::COL_TASKS; ++t) fftc_col<SCfg>(b, t);
        fft_phase(FFT_PH_O);
    }
    double max_err = 0, max_ref = 0;
    for (int p = 0; p < Cfg::G; ++p)
        for (int i = 0; i < Cfg::HO; ++i)
            for (int j = 0; j < Cfg::WO; ++j) {
                double acc = 0;
                for (int u = 0; u < Cfg::KH; ++u) {
                    int r = i + u - Cfg::PH;
                    if (Cfg::CIRC) r = ((r % Cfg::HX) + Cfg::HX) % Cfg::HX;
                    for (int v = 0; v < Cfg::KW; ++v) {
                        int c = j + v - Cfg::PW;
                        c = c < 0 ? 0 : (c > Cfg::WX - 1 ? Cfg::WX - 1 : c);
                        acc += (double)x[p * Cfg::XPL + r * Cfg::WX + c] * k[p * Cfg::KPL + u * Cfg::KW + v];
                    }
                }
                const double e = std::fabs(acc - out[p * Cfg::OPL + i * Cfg::WO + j]);
                if (!(e <= max_err)) max_err = e;
                if (std::fabs(acc) > max_ref) max_ref = std::fabs(acc);
            }
    printf("%s max_err %.3e max_ref %.3e order_free 1\n", name, max_err, max_ref);
    return max_err <= 2e-5 * max_ref ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += check<FCfg<29, 29, 61, 61, false, 2, 192>>("k1_256", 1);
    bad += check<FCfg<29, 29, 29, 29, true, 2, 128>>("k2_256", 2);
    bad += check<FCfg<15, 15, 39, 39, false, 2, 128>>("win15", 3);
    bad += check<FCfg<13, 11, 40, 37, false, 4, 256>>("ragged", 4);
    bad += check<FCfg<12, 9, 20, 21, true, 4, 256>>("ragged_circ", 5);
    bad += check_spectra<FCfg<29, 29, 61, 61, false, 2, 192>, FCfg<29, 29, 61, 61, false, 2, 192, HDN_FFT_TB, true>>("k1_256_spectra", 6);
    bad += check_spectra<FCfg<29, 29, 29, 29, true, 2, 128>, FCfg<29, 29, 29, 29, true, 2, 128, HDN_FFT_TB, true>>("k2_256_spectra", 7);
    bad += check_spectra<FCfg<29, 29, 61, 61, false, 2, 192>, FCfg<29, 29, 61, 61, false, 2, 224, HDN_FFT_TB, true>>("k1_256_spectra_pipe", 8);
    return bad ? 1 : 0;
}
