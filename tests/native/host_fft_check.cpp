// host_fft_check.cpp -- CPU check of the FFT correlation arithmetic (hdn_b200/csrc/xcorr_fft.cuh, fft64.cuh).
// Runs the kernel's three phases task by task (the order the barriers of xcorr_fft_kernel impose) on one group of planes
// and compares with a direct double-precision correlation.  Test infrastructure only: built and run by
// tests/test_fft_host.py with g++; prints "<config> max_err <e> max_ref <m>" per shape, exit code 1 if any error exceeds 2e-5*max_ref.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../hdn_b200/csrc/xcorr_fft.cuh"

using namespace hdn;

template <class Cfg>
static int check(const char *name, unsigned seed) {
    std::vector<float> raw(Cfg::RAW_FLOATS), out(Cfg::OUT_FLOATS, -1.f);
    std::vector<float2> XR(Cfg::G * Cfg::XR_PLANE), KR(Cfg::G * Cfg::KR_PLANE);
    for (auto &v : XR) v = float2{NAN, NAN};  // anything read before it is written poisons the result
    for (auto &v : KR) v = float2{NAN, NAN};
    srand(seed);
    for (auto &v : raw) v = (float)rand() / RAND_MAX * 2.f - 0.7f;  // non-zero mean, like post-ReLU features
    const float *rawx = raw.data(), *rawk = raw.data() + Cfg::G * Cfg::XPL;
    for (int s = 0; s < Cfg::R_SLOTS; ++s) fftc_phase_R<Cfg>(rawx, rawk, XR.data(), KR.data(), s);
    for (int t = 0; t < Cfg::C_TASKS; ++t) fftc_phase_C<Cfg>(XR.data(), KR.data(), t);
    for (int t = 0; t < Cfg::O_TASKS; ++t) fftc_phase_O<Cfg>(XR.data(), out.data(), t);
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
    printf("%s max_err %.3e max_ref %.3e\n", name, max_err, max_ref);
    return (max_err <= 2e-5 * max_ref) ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += check<FCfg<29, 29, 61, 61, false, 4, 256>>("k1_256", 1);
    bad += check<FCfg<29, 29, 29, 29, true, 4, 256>>("k2_256", 2);
    bad += check<FCfg<15, 15, 39, 39, false, 4, 256>>("win15", 3);
    bad += check<FCfg<13, 11, 40, 37, false, 4, 256>>("ragged", 4);
    return bad ? 1 : 0;
}
