// Host emulation of the device phase functions in diffmusic_b200/csrc/{fft_core,stft_frame}.cuh.
// TEST INFRASTRUCTURE: compiled with g++ by tests/test_cpu_emulation.py; never shipped, never on the product path.
// Every phase is run for all 64 "threads" of a frame group before the next phase starts (= the GPU barrier).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../diffmusic_b200/csrc/stft_frame.cuh"

using namespace dm;

extern "C" {

// complex FFT of size 512 / 4096 through the Stockham passes (in: interleaved re,im; out likewise)
void emul_fft(int n, int inverse, const float* tw, const float* in, float* out) {
    std::vector<float> are(swz_len(n)), aim(swz_len(n)), bre(swz_len(n)), bim(swz_len(n));
    const cf* t = reinterpret_cast<const cf*>(tw);
    for (int i = 0; i < n; ++i) { are[swz(i)] = in[2 * i]; aim[swz(i)] = in[2 * i + 1]; }
    SwzLoad la{are.data(), aim.data()}, lb{bre.data(), bim.data()};
    SwzStore sa{are.data(), aim.data()}, sb{bre.data(), bim.data()};
    const float* rr; const float* ri;
    if (n == 512) {
        for (int j = 0; j < 64; ++j) inverse ? stockham_pass<512, 1, 1>(j, t, la, sb) : stockham_pass<512, 1, -1>(j, t, la, sb);
        for (int j = 0; j < 64; ++j) inverse ? stockham_pass<512, 8, 1>(j, t, lb, sa) : stockham_pass<512, 8, -1>(j, t, lb, sa);
        for (int j = 0; j < 64; ++j) inverse ? stockham_pass<512, 64, 1>(j, t, la, sb) : stockham_pass<512, 64, -1>(j, t, la, sb);
        rr = bre.data(); ri = bim.data();
    } else {
        for (int j = 0; j < 512; ++j) inverse ? stockham_pass<4096, 1, 1>(j, t, la, sb) : stockham_pass<4096, 1, -1>(j, t, la, sb);
        for (int j = 0; j < 512; ++j) inverse ? stockham_pass<4096, 8, 1>(j, t, lb, sa) : stockham_pass<4096, 8, -1>(j, t, lb, sa);
        for (int j = 0; j < 512; ++j) inverse ? stockham_pass<4096, 64, 1>(j, t, la, sb) : stockham_pass<4096, 64, -1>(j, t, la, sb);
        for (int j = 0; j < 512; ++j) inverse ? stockham_pass<4096, 512, 1>(j, t, lb, sa) : stockham_pass<4096, 512, -1>(j, t, lb, sa);
        rr = are.data(); ri = aim.data();
    }
    for (int i = 0; i < n; ++i) { out[2 * i] = rr[swz(i)]; out[2 * i + 1] = ri[swz(i)]; }
}

}  // extern "C"

struct EmulTables {
    const float* window; const float* tw512; const float* w1024;
    const int* mel_kstart; const int* mel_klen; const float* mel_w; int mel_wstride;
    const int* bin_m0; const float* bin_w0; const float* bin_w1;
};

template <int MODE>
static void run(const EmulTables& e, int clamp, const float* y, long long Ly, int hop, const float* mask,
                const float* ref, float* out, float* ypbar, double* sumsq) {
    StftTables t{e.window, reinterpret_cast<const cf*>(e.tw512), reinterpret_cast<const cf*>(e.w1024),
                 e.mel_kstart, e.mel_klen, e.mel_w, e.mel_wstride, e.bin_m0, e.bin_w0, e.bin_w1};
    const long long T = 1 + Ly / hop;
    const int nrow = (MODE == kModePhaseWav) ? kBins : kMels;
    std::vector<float> buf(kFrameSmemFloats, 0.f), frame(kNfft);
    float* p = buf.data();
    FrameSmem s;
    s.a_re = p; p += swz_len(kH); s.a_im = p; p += swz_len(kH);
    s.b_re = p; p += swz_len(kH); s.b_im = p; p += swz_len(kH);
    s.melbar = p; p += 72; s.aux = p;
    if (ypbar) std::memset(ypbar, 0, sizeof(float) * (Ly + 1024));
    std::vector<ThreadConsts> tc(64);
    for (int tid = 0; tid < 64; ++tid) load_thread_consts(tid, t, tc[tid]);
    std::vector<float> melw_t(e.mel_w, e.mel_w + (size_t)e.mel_wstride * kMels);  // already [i][64]
    double acc = 0.0;
    for (long long f = 0; f < T; ++f) {
        for (int n = 0; n < kNfft; ++n) {
            long long j = reflect_src(f * hop + n, Ly);
            frame[n] = y[j] * (mask ? mask[j] : 1.f);
        }
        for (int tid = 0; tid < 64; ++tid) fwd_pass1<false>(tid, frame.data(), t.window, s);
        for (int tid = 0; tid < 64; ++tid) fwd_pass2(tid, tc[tid], s);
        for (int tid = 0; tid < 64; ++tid) fwd_pass3(tid, tc[tid], s);
        for (int tid = 0; tid < 64; ++tid) fwd_unpack<MODE>(tid, t.w1024, s);
        if (MODE == kModePhaseWav) {
            for (int k = 0; k < kBins; ++k) {
                float mag = p_at(s, k);
                if (out) out[k * T + f] = mag;
                if (ref) { float d = ref[k * T + f] - mag; acc += (double)d * d; p_at(s, k) = -d; }
            }
        } else {
            for (int m = 0; m < 64; ++m) {
                float v;
                float d2 = mel_residual<MODE>(m, tc[m], melw_t.data(), s, clamp != 0, ref != nullptr,
                                              ref ? ref[m * T + f] : 0.f, &v);
                acc += d2;
                if (out) out[m * T + f] = v;
            }
        }
        if (!ypbar) continue;
        for (int tid = 0; tid < 64; ++tid) bwd_pack<MODE>(tid, t, s);
        for (int tid = 0; tid < 64; ++tid) inv_pass1(tid, s);
        for (int tid = 0; tid < 64; ++tid) inv_pass2(tid, tc[tid], s);
        for (int tid = 0; tid < 64; ++tid) inv_pass3(tid, tc[tid], s);
        for (int n = 0; n < kNfft; ++n) ypbar[f * hop + n] += frame_grad_sample(s, n) * t.window[n];
    }
    (void)nrow;
    *sumsq = acc;
}

extern "C" {
// y: (Ly,), ref/out: (64 or 513, T) row-major, ypbar: (Ly + 1024,) unscaled padded gradient (d(sum d^2)/2 ... see test)
void emul_stft_guidance(const EmulTables* e, int mode, int clamp, const float* y, long long Ly, int hop,
                        const float* mask, const float* ref, float* out, float* ypbar, double* sumsq) {
    if (mode == 0) run<kModeMelDb>(*e, clamp, y, Ly, hop, mask, ref, out, ypbar, sumsq);
    else if (mode == 1) run<kModePhaseMel>(*e, clamp, y, Ly, hop, mask, ref, out, ypbar, sumsq);
    else run<kModePhaseWav>(*e, clamp, y, Ly, hop, mask, ref, out, ypbar, sumsq);
}
}

// ------------------------------------------------------------------------------------------------------------------
// Frame-pair pipeline (diffmusic_b200/csrc/stft_pair.cuh): two frames per 64-"thread" group, per-thread spectrum
// registers carried from the unpack phase to the pack phase, gathered overlap-add as in stft_pair_kernel.
#include "../../diffmusic_b200/csrc/stft_pair.cuh"

template <int MODE>
static void run_pair(const EmulTables& e, int clamp, const float* y, long long Ly, int hop, const float* mask,
                     const float* ref, float* out, float* ypbar, double* sumsq) {
    StftTables t{e.window, reinterpret_cast<const cf*>(e.tw512), reinterpret_cast<const cf*>(e.w1024),
                 e.mel_kstart, e.mel_klen, e.mel_w, e.mel_wstride, e.bin_m0, e.bin_w0, e.bin_w1};
    const long long T = 1 + Ly / hop;
    const int hop2 = hop / 2;
    std::vector<float> buf(kPairSmemFloats + 8, 0.f), fa(kNfft), fb(kNfft);
    // 16-byte aligned carve-up
    float* base = buf.data();
    while (reinterpret_cast<uintptr_t>(base) & 15) ++base;
    PairSmem s;
    s.a = reinterpret_cast<c2*>(base);
    s.b = reinterpret_cast<c2*>(base + 4 * kH);
    if (ypbar) std::memset(ypbar, 0, sizeof(float) * (Ly + 1024));
    std::vector<PairConsts> pc(64);
    std::vector<PairX> px(64);
    struct Regs { cf v[8]; };
    std::vector<Regs> rva(64), rvb(64);
    std::vector<float> accbuf_raw(Ly + 1024 + 8, 0.f);
    struct AccView { float* p; float* data() { return p; } } accbuf{accbuf_raw.data()};
    while (reinterpret_cast<uintptr_t>(accbuf.p) & 7) ++accbuf.p;
    std::vector<f2> binw(kBins);
    std::vector<unsigned char> binm(kBins);
    for (int k = 0; k < kBins; ++k) { binw[k] = f2{e.bin_w0[k], e.bin_w1[k]}; binm[k] = (unsigned char)e.bin_m0[k]; }
    const PairBinTab bins{binw.data(), binm.data()};
    for (int tid = 0; tid < 64; ++tid) load_pair_consts(tid, t, pc[tid]);
    double acc = 0.0;
    for (long long f = 0; f < T; f += 2) {
        const bool has_b = f + 1 < T;
        const long long f2i = has_b ? f + 1 : f;
        for (int n = 0; n < kNfft; ++n) {
            long long ja = reflect_src(f * hop + n, Ly), jb = reflect_src(f2i * hop + n, Ly);
            fa[n] = y[ja] * (mask ? mask[ja] : 1.f);
            fb[n] = y[jb] * (mask ? mask[jb] : 1.f);
        }
        for (int tid = 0; tid < 64; ++tid) pair_fwd_pass1(tid, fa.data(), fb.data(), t.window, s);
        for (int tid = 0; tid < 64; ++tid) pair_fwd_pass2(tid, pc[tid], s);
        for (int tid = 0; tid < 64; ++tid) pair_fwd_pass3(tid, pc[tid], s);
        for (int tid = 0; tid < 64; ++tid) pair_unpack<MODE>(tid, pc[tid], s, px[tid]);
        f2* P = pair_energy(s);
        if (MODE == kModePhaseWav) {
            for (int k = 0; k < kBins; ++k) {
                f2 mag = P[k];
                if (out) { out[k * T + f] = mag.x; if (has_b) out[k * T + f2i] = mag.y; }
                if (ref) {
                    float da = ref[k * T + f] - mag.x, db = ref[k * T + f2i] - mag.y;
                    acc += (double)da * da;
                    if (has_b) acc += (double)db * db;
                    P[k] = f2{-da, -db};
                }
            }
        } else {
            std::vector<f2> own(64);
            for (int m = 0; m < 64; ++m) own[m] = pair_mel_project(m, pc[m], e.mel_w, s);
            for (int m = 0; m < 64; ++m) {
                f2 mel = pair_mel_combine(m, own[m], s);
                float va, da, vb, db;
                mel_value<MODE>(mel.x, clamp != 0, va, da);
                mel_value<MODE>(mel.y, clamp != 0, vb, db);
                if (ref) {
                    float ra = ref[m * T + f] - va, rb = ref[m * T + f2i] - vb;
                    pair_melbar(s)[m] = f2{-ra * da, -rb * db};
                    acc += (double)ra * ra;
                    if (has_b) acc += (double)rb * rb;
                }
                if (out) { out[m * T + f] = va; if (has_b) out[m * T + f2i] = vb; }
            }
        }
        if (!ypbar) continue;
        for (int tid = 0; tid < 64; ++tid) pair_pack<MODE>(tid, pc[tid], bins, s, px[tid]);
        for (int tid = 0; tid < 64; ++tid) pair_inv_pass1(tid, s);
        for (int tid = 0; tid < 64; ++tid) pair_inv_pass2(tid, pc[tid], s);
        // last pass into "registers", then the ordered overlap-add from them: frame A, (barrier), frame B
        for (int tid = 0; tid < 64; ++tid) pair_inv_pass3(tid, pc[tid], s, rva[tid].v, rvb[tid].v);
        f2* acc2 = reinterpret_cast<f2*>(accbuf.data());
        for (int tid = 0; tid < 64; ++tid) pair_ola_add(tid, t.window, rva[tid].v, acc2 + f * hop2);
        if (has_b)
            for (int tid = 0; tid < 64; ++tid) pair_ola_add(tid, t.window, rvb[tid].v, acc2 + f2i * hop2);
    }
    if (ypbar) std::memcpy(ypbar, accbuf.data(), sizeof(float) * (Ly + 1024));
    *sumsq = acc;
}

extern "C" {
void emul_stft_guidance_pair(const EmulTables* e, int mode, int clamp, const float* y, long long Ly, int hop,
                             const float* mask, const float* ref, float* out, float* ypbar, double* sumsq) {
    if (mode == 0) run_pair<kModeMelDb>(*e, clamp, y, Ly, hop, mask, ref, out, ypbar, sumsq);
    else if (mode == 1) run_pair<kModePhaseMel>(*e, clamp, y, Ly, hop, mask, ref, out, ypbar, sumsq);
    else run_pair<kModePhaseWav>(*e, clamp, y, Ly, hop, mask, ref, out, ypbar, sumsq);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Warp-per-frame-pair pipeline (diffmusic_b200/csrc/stft_warp.cuh): 32 "lanes" carry two frames as one 32 x 32 complex
// FFT; a shuffle is a phase boundary (all lanes publish, then all lanes read), the gathered overlap-add as in
// stft_warp_kernel.
#include "../../diffmusic_b200/csrc/stft_warp.cuh"

template <int MODE>
static void run_warp(const EmulTables& e, int clamp, const float* y, long long Ly, int hop, const float* mask,
                     const float* ref, float* out, float* ypbar, double* sumsq, const float* image, int na, int nb,
                     int pair_shift) {
    const long long T = 1 + Ly / hop;
    std::vector<float> raw(kWarpBufFloats + 8, 0.f), fa(kNfft), fb(kNfft);
    float* wbuf = raw.data();
    while (reinterpret_cast<uintptr_t>(wbuf) & 15) ++wbuf;
    cf* xbuf = reinterpret_cast<cf*>(wbuf);
    f2* P = reinterpret_cast<f2*>(wbuf + kWarpPOff);
    f2* melbar = reinterpret_cast<f2*>(wbuf + kWarpMelbarOff);
    // the host-built table image (diffmusic_b200/tables.py warp_image), 16-byte aligned like its shared-memory copy
    const WarpImage il = warp_image_layout(na, nb);
    std::vector<float> imgraw(il.total + 8);
    float* img = imgraw.data();
    while (reinterpret_cast<uintptr_t>(img) & 15) ++img;
    std::memcpy(img, image, sizeof(float) * il.total);
    const f2* win2 = reinterpret_cast<const f2*>(img + il.win2);
    const f4* tw4 = reinterpret_cast<const f4*>(img + il.tw4);
    const f2* melp = reinterpret_cast<const f2*>(img + il.melp);
    const int* lanek = reinterpret_cast<const int*>(img + il.lanek);
    const PairBinTab bins{reinterpret_cast<const f2*>(img + il.binw), reinterpret_cast<const unsigned char*>(img + il.binm)};
    struct Lane { cf v[32]; WarpX x; cf snd, rcv, prev; };
    std::vector<Lane> L(32);
    if (ypbar) std::memset(ypbar, 0, sizeof(float) * (Ly + 1024));
    double acc = 0.0;
    // pair_shift = 1: frame 0 rides alone (as the last frame of an odd tile does), the pairs are (1, 2), (3, 4), ...
    for (long long f = 0; f < T; f += (pair_shift && f == 0) ? 1 : 2) {
        const bool has_b = f + 1 < T && !(pair_shift && f == 0);
        const long long f2i = has_b ? f + 1 : f;
        for (int n = 0; n < kNfft; ++n) {
            long long ja = reflect_src(f * hop + n, Ly), jb = reflect_src(f2i * hop + n, Ly);
            fa[n] = y[ja] * (mask ? mask[ja] : 1.f);
            fb[n] = y[jb] * (mask ? mask[jb] : 1.f);
        }
        float ssa = 0.f, ssb = 0.f;
        for (int l = 0; l < 32; ++l) {
            float sa, sb;
            warp_load_frames(l, fa.data(), fb.data(), win2, L[l].v, sa, sb);
            ssa += sa;
            ssb += sb;
        }
        float bal_s, bal_inv;
        bool zero_a, zero_b;
        warp_balance(ssa, ssb, bal_s, bal_inv, zero_a, zero_b);
        for (int l = 0; l < 32; ++l) {
            warp_scale_b(L[l].v, bal_s);
            dft32<-1>(L[l].v);
            warp_twiddle_store<-1>(l, L[l].v, tw4, xbuf);
        }
        for (int l = 0; l < 32; ++l) {
            warp_xchg_load(l, xbuf, L[l].v);
            dft32<-1>(L[l].v);
        }
        if (MODE != kModePhaseWav) {
            for (int l = 0; l < 8; ++l) melbar[64 + l] = f2{0.f, 0.f};
            P[kH + 1] = f2{0.f, 0.f};
        }
        // split, one owned bin per step (a shuffle = all lanes publish, then all lanes read)
        for (int i = 0; i < 17; ++i) {
            if (i < 16) {
                for (int l = 0; l < 32; ++l) L[l].snd = warp_split_send_i(l, L[l].v, i);
                for (int l = 0; l < 32; ++l) L[l].rcv = L[(32 - l) & 31].snd;
            }
            for (int l = 0; l < 32; ++l) {
                WarpX& x = L[l].x;
                if (i < 16) warp_split_recv_i(L[l].v, i, L[l].rcv, x.a[i], x.b[i]);
                else warp_split_nyquist(L[l].v, x.a[16], x.b[16]);
                x.b[i] = cf{x.b[i].x * bal_inv, x.b[i].y * bal_inv};
                if (zero_a) x.a[i] = cf{0.f, 0.f};
                if (zero_b) x.b[i] = cf{0.f, 0.f};
                const bool mine = i < 16 || l == 0;
                const int k = warp_bin_of(l, i);
                const f2 en = warp_bin_energies<MODE>(x.a[i], x.b[i]);
                if (MODE == kModePhaseWav) {
                    f2 gk = f2{0.f, 0.f};
                    if (mine) {
                        if (out) { out[k * T + f] = en.x; if (has_b) out[k * T + f2i] = en.y; }
                        if (ref) {
                            float da = ref[k * T + f] - en.x, db = ref[k * T + f2i] - en.y;
                            acc += (double)da * da;
                            if (has_b) acc += (double)db * db;
                            gk = f2{-da, -db};
                        }
                    }
                    x.a[i] = warp_xbar<MODE>(x.a[i], gk.x);
                    x.b[i] = warp_xbar<MODE>(x.b[i], gk.y);
                } else if (mine) {
                    P[k] = en;
                }
            }
        }
        if (MODE != kModePhaseWav) {
            std::vector<f2> lo(32), hi(32);
            for (int l = 0; l < 32; ++l) warp_mel_project(l, na, nb, lanek[l], lanek[32 + l], melp, P, lo[l], hi[l]);
            for (int l = 0; l < 32; ++l) {
                const int ms[2] = {lanek[64 + l], lanek[96 + l]};
                const f2 mel[2] = {lo[l], hi[l]};
                for (int q = 0; q < 2; ++q) {
                    const int m = ms[q];
                    float va, da, vb, db;
                    mel_value<MODE>(mel[q].x, clamp != 0, va, da);
                    mel_value<MODE>(mel[q].y, clamp != 0, vb, db);
                    if (ref) {
                        float ra = ref[m * T + f] - va, rb = ref[m * T + f2i] - vb;
                        melbar[m] = f2{-ra * da, -rb * db};
                        acc += (double)ra * ra;
                        if (has_b) acc += (double)rb * rb;
                    }
                    if (out) { out[m * T + f] = va; if (has_b) out[m * T + f2i] = vb; }
                }
            }
        }
        if (!ypbar) continue;
        // backward, one owned bin per step
        auto xbar = [&](int l, int i, cf& ya, cf& yb) {
            ya = L[l].x.a[i];
            yb = L[l].x.b[i];
            if (MODE != kModePhaseWav) {
                const f2 gk = warp_bin_cotangent(warp_bin_of(l, i), bins, melbar);
                ya = warp_xbar<MODE>(ya, gk.x);
                yb = warp_xbar<MODE>(yb, gk.y);
            }
        };
        for (int i = 0; i < 16; ++i) {
            for (int l = 0; l < 32; ++l) {
                cf ya, yb, q, qm;
                xbar(l, i, ya, yb);
                warp_q_pair(ya, yb, q, qm);
                if (i == 0 && l == 0) q = warp_q_real(ya, yb);
                L[l].v[i] = q;
                L[l].snd = qm;
            }
            for (int l = 0; l < 32; ++l) L[l].rcv = L[(32 - l) & 31].snd;
            for (int l = 0; l < 32; ++l) {
                if (i > 0) L[l].v[32 - i] = l == 0 ? L[l].rcv : L[l].prev;
                L[l].prev = L[l].rcv;
            }
        }
        for (int l = 0; l < 32; ++l) {
            cf ya, yb;
            xbar(l, 16, ya, yb);
            L[l].v[16] = l == 0 ? warp_q_real(ya, yb) : L[l].prev;
            dft32<+1>(L[l].v);
        }
        for (int l = 0; l < 32; ++l) warp_twiddle_store<+1>(l, L[l].v, tw4, xbuf);
        for (int l = 0; l < 32; ++l) {
            warp_xchg_load(l, xbuf, L[l].v);
            dft32<+1>(L[l].v);
        }
        for (int l = 0; l < 32; ++l) warp_store_gradients(l, L[l].v, win2, wbuf);
        for (int n = 0; n < kNfft; ++n) {
            ypbar[f * hop + n] += wbuf[n];
            if (has_b) ypbar[f2i * hop + n] += wbuf[kWarpGStride + n];
        }
    }
    *sumsq = acc;
}

extern "C" {
void emul_stft_guidance_warp(const EmulTables* e, int mode, int clamp, const float* y, long long Ly, int hop,
                             const float* mask, const float* ref, float* out, float* ypbar, double* sumsq,
                             const float* image, int na, int nb, int pair_shift) {
    if (mode == 0) run_warp<kModeMelDb>(*e, clamp, y, Ly, hop, mask, ref, out, ypbar, sumsq, image, na, nb, pair_shift);
    else if (mode == 1) run_warp<kModePhaseMel>(*e, clamp, y, Ly, hop, mask, ref, out, ypbar, sumsq, image, na, nb, pair_shift);
    else run_warp<kModePhaseWav>(*e, clamp, y, Ly, hop, mask, ref, out, ypbar, sumsq, image, na, nb, pair_shift);
}

// LogSpectralDistance frames (csrc/metrics.cu lsd_pair_kernel): frame t of the reference clip rides as frame A, frame t of
// the estimate as frame B of ONE pair FFT; per-frame distance from the magnitudes side by side in P.
void emul_lsd_frames(const EmulTables* e, const float* ref, const float* est, long long L, int hop, int pad_reflect,
                     float eps, float* out) {
    StftTables t{e->window, reinterpret_cast<const cf*>(e->tw512), reinterpret_cast<const cf*>(e->w1024),
                 e->mel_kstart, e->mel_klen, e->mel_w, e->mel_wstride, e->bin_m0, e->bin_w0, e->bin_w1};
    const long long T = 1 + L / hop;
    std::vector<float> buf(kPairSmemFloats + 8, 0.f), fa(kNfft), fb(kNfft);
    float* base = buf.data();
    while (reinterpret_cast<uintptr_t>(base) & 15) ++base;
    PairSmem s;
    s.a = reinterpret_cast<c2*>(base);
    s.b = reinterpret_cast<c2*>(base + 4 * kH);
    std::vector<PairConsts> pc(64);
    std::vector<PairX> px(64);
    for (int tid = 0; tid < 64; ++tid) load_pair_consts(tid, t, pc[tid]);
    auto sanitize = [](float v) { return v != v ? 0.f : (std::isinf(v) ? (v > 0 ? 1.f : -1.f) : v); };
    for (long long f = 0; f < T; ++f) {
        for (int n = 0; n < kNfft; ++n) {
            long long j = f * hop + n - kNfft / 2;
            bool inside = j >= 0 && j < L;
            if (!inside && pad_reflect) { j = reflect_src(f * hop + n, L); inside = true; }
            fa[n] = inside ? ref[j] : 0.f;
            fb[n] = inside ? sanitize(est[j]) : 0.f;
        }
        for (int tid = 0; tid < 64; ++tid) pair_fwd_pass1(tid, fa.data(), fb.data(), t.window, s);
        for (int tid = 0; tid < 64; ++tid) pair_fwd_pass2(tid, pc[tid], s);
        for (int tid = 0; tid < 64; ++tid) pair_fwd_pass3(tid, pc[tid], s);
        for (int tid = 0; tid < 64; ++tid) pair_unpack<kModePhaseWav>(tid, pc[tid], s, px[tid]);
        const f2* P = pair_energy(s);
        float acc = 0.f;
        for (int k = 0; k < kBins; ++k) {
            const float d = log10f(P[k].x + eps) - log10f(P[k].y + eps);
            acc += d * d;
        }
        out[f] = sqrtf(acc / kBins);
    }
}

// mel_spectrogram_to_waveform_with_phase (csrc/istft.cu): relu(W mel) * exp(i phase) per bin -> pair-major spectrum cells ->
// pair_pack_spectrum -> the three inverse passes -> rectangular overlap-add -> envelope division and trimming.
// mel: (64, T), phase: (513, T), winv_t: (64, 513), out: (out_len).
void emul_istft_mel_phase(const EmulTables* e, const float* winv_t, const float* mel, const float* phase, long long T,
                          int hop, float* out, long long out_len) {
    StftTables t{e->window, reinterpret_cast<const cf*>(e->tw512), reinterpret_cast<const cf*>(e->w1024),
                 e->mel_kstart, e->mel_klen, e->mel_w, e->mel_wstride, e->bin_m0, e->bin_w0, e->bin_w1};
    std::vector<float> buf(kPairSmemFloats + 8, 0.f), cells(4 * 516 + 8, 0.f);
    float* base = buf.data();
    while (reinterpret_cast<uintptr_t>(base) & 15) ++base;
    float* cb = cells.data();
    while (reinterpret_cast<uintptr_t>(cb) & 15) ++cb;
    PairSmem s;
    s.a = reinterpret_cast<c2*>(base);
    s.b = reinterpret_cast<c2*>(base + 4 * kH);
    c2* spec = reinterpret_cast<c2*>(cb);
    std::vector<PairConsts> pc(64);
    std::vector<PairX> px(64);
    for (int tid = 0; tid < 64; ++tid) load_pair_consts(tid, t, pc[tid]);
    const long long ola_len = (long long)hop * (T - 1) + kNfft;
    std::vector<float> ola(ola_len, 0.f);
    for (long long fa = 0; fa < T; fa += 2) {
        const bool has_b = fa + 1 < T;
        for (int k = 0; k < kBins; ++k) {
            float lin[2] = {0.f, 0.f};
            for (int m = 0; m < kMels; ++m) {
                lin[0] = fmaf(winv_t[m * kBins + k], mel[m * T + fa], lin[0]);
                if (has_b) lin[1] = fmaf(winv_t[m * kBins + k], mel[m * T + fa + 1], lin[1]);
            }
            float x[4] = {0.f, 0.f, 0.f, 0.f};
            for (int f = 0; f < (has_b ? 2 : 1); ++f) {
                const float mag = lin[f] < 0.f ? 0.f : lin[f];
                const float ph = phase[k * T + fa + f];
                x[2 * f] = mag * cosf(ph);
                x[2 * f + 1] = mag * sinf(ph);
            }
            spec[k] = c2{x[0], x[1], x[2], x[3]};
        }
        for (int tid = 0; tid < 64; ++tid) pair_load_spectrum(tid, spec, px[tid]);
        for (int tid = 0; tid < 64; ++tid) pair_pack_spectrum(tid, pc[tid], s, px[tid]);
        for (int tid = 0; tid < 64; ++tid) pair_inv_pass1(tid, s);
        for (int tid = 0; tid < 64; ++tid) pair_inv_pass2(tid, pc[tid], s);
        std::vector<cf> va(64 * 8), vb(64 * 8);
        for (int tid = 0; tid < 64; ++tid) {
            cf a[8], b[8];
            pair_inv_pass3(tid, pc[tid], s, a, b);
            for (int q = 0; q < 8; ++q) { va[tid * 8 + q] = a[q]; vb[tid * 8 + q] = b[q]; }
        }
        for (int f = 0; f < (has_b ? 2 : 1); ++f)
            for (int tid = 0; tid < 64; ++tid) {
                cf v[8];
                for (int q = 0; q < 8; ++q) v[q] = f ? vb[tid * 8 + q] : va[tid * 8 + q];
                pair_ola_add_rect(tid, v, reinterpret_cast<f2*>(ola.data() + (fa + f) * hop));
            }
    }
    const long long n_valid = (long long)hop * (T - 1);
    for (long long j = 0; j < out_len; ++j) {
        float v = 0.f;
        if (j < n_valid) {
            const long long i = j + kNfft / 2;
            const long long t_hi = std::min(T - 1, i / hop);
            const long long t_lo = i >= kNfft ? (i - kNfft) / hop + 1 : 0;
            v = ola[i] * (1.0f / kNfft) / (float)(t_hi - t_lo + 1);
        }
        out[j] = v;
    }
}

// waveform_to_spectrogram (csrc/istft.cu spectrogram_pair_kernel): frames t, t + 1 through one pair FFT, magnitude and
// atan2 of the owned bins from the registers.  mag / phase: (513, T).
void emul_stft_spectrogram(const EmulTables* e, const float* wav, long long L, int hop, float* mag, float* phase) {
    StftTables t{e->window, reinterpret_cast<const cf*>(e->tw512), reinterpret_cast<const cf*>(e->w1024),
                 e->mel_kstart, e->mel_klen, e->mel_w, e->mel_wstride, e->bin_m0, e->bin_w0, e->bin_w1};
    const long long T = 1 + L / hop;
    std::vector<float> buf(kPairSmemFloats + 8, 0.f), fa(kNfft), fb(kNfft);
    float* base = buf.data();
    while (reinterpret_cast<uintptr_t>(base) & 15) ++base;
    PairSmem s;
    s.a = reinterpret_cast<c2*>(base);
    s.b = reinterpret_cast<c2*>(base + 4 * kH);
    std::vector<PairConsts> pc(64);
    std::vector<PairX> px(64);
    for (int tid = 0; tid < 64; ++tid) load_pair_consts(tid, t, pc[tid]);
    auto emit = [&](int k, long long ta, bool has_b, cf xa, cf xb) {
        mag[k * T + ta] = pair_bin_energy<kModePhaseWav>(xa);
        phase[k * T + ta] = atan2f(xa.y, xa.x);
        if (has_b) {
            mag[k * T + ta + 1] = pair_bin_energy<kModePhaseWav>(xb);
            phase[k * T + ta + 1] = atan2f(xb.y, xb.x);
        }
    };
    for (long long f = 0; f < T; f += 2) {
        const bool has_b = f + 1 < T;
        for (int n = 0; n < kNfft; ++n) {
            fa[n] = wav[reflect_src(f * hop + n, L)];
            fb[n] = wav[reflect_src((has_b ? f + 1 : f) * hop + n, L)];
        }
        for (int tid = 0; tid < 64; ++tid) pair_fwd_pass1(tid, fa.data(), fb.data(), t.window, s);
        for (int tid = 0; tid < 64; ++tid) pair_fwd_pass2(tid, pc[tid], s);
        for (int tid = 0; tid < 64; ++tid) pair_fwd_pass3(tid, pc[tid], s);
        for (int tid = 0; tid < 64; ++tid) pair_unpack<kModePhaseWav>(tid, pc[tid], s, px[tid]);
        for (int tid = 0; tid < 64; ++tid) {
            for (int i = 0; i < 4; ++i) {
                const int k = tid + 64 * i;
                emit(k, f, has_b, px[tid].lo[i][0], px[tid].lo[i][1]);
                emit(kH - k, f, has_b, px[tid].hi[i][0], px[tid].hi[i][1]);
            }
            if (tid == 0) emit(kH / 2, f, has_b, px[0].q[0], px[0].q[1]);
        }
    }
}

// Bank-conflict audit of the cell swizzle: a 128-bit shared access is served per quarter-warp (8 consecutive lanes),
// conflict-free iff the 8 cells fall into 8 distinct 16-byte bank groups (cell index mod 8).  Returns the number of
// (pattern, quarter-warp) instances with a conflict, and checks st_c2_addr / the load form against sw4 of the logical index.
int emul_check_pair_swizzle() {
    int bad = 0;
    auto distinct8 = [](const int* cells) {
        int seen = 0;
        for (int l = 0; l < 8; ++l) seen |= 1 << (cells[l] & 7);
        return seen == 0xff;
    };
    for (int q0 = 0; q0 < 64; q0 += 8)        // quarter-warp = lanes q0 .. q0+7
        for (int r = 0; r < 8; ++r) {
            int ld[8], s1[8], s8[8], s64[8], ulo[8], uhi[8];
            for (int l = 0; l < 8; ++l) {
                const int j = q0 + l;
                ld[l] = sw4(j) + 64 * r;
                bad += ld[l] != sw4(j + 64 * r);
                s1[l] = st_c2_addr<1>(j, r);
                bad += s1[l] != sw4(8 * j + r);
                s8[l] = st_c2_addr<8>(j, r);
                bad += s8[l] != sw4(64 * (j >> 3) + (j & 7) + 8 * r);
                s64[l] = st_c2_addr<64>(j, r);
                bad += s64[l] != sw4(j + 64 * r);
                if (r < 4) {
                    const int k = j + 64 * r;
                    ulo[l] = sw4(k);
                    uhi[l] = sw4((kH - k) & (kH - 1));
                    if (j >= 1) bad += (sw4((kH - j) & (kH - 1)) - 64 * r) != sw4(kH - k);
                }
            }
            bad += !distinct8(ld) + !distinct8(s1) + !distinct8(s8) + !distinct8(s64);
            if (r < 4) bad += !distinct8(ulo);
            (void)uhi;  // the descending side may see one 2-way conflict where a quarter-warp straddles a multiple of 8
        }
    return bad;
}
}

// ------------------------------------------------------------------------------------------------------------------
// Overlap-save RIR correlation / adjoint (diffmusic_b200/csrc/rir_block.cuh), same block loop as rir_conv.cu
#include "../../diffmusic_b200/csrc/rir_block.cuh"

namespace {
struct EmulRir {
    std::vector<float> buf;
    RirSmem s;
    EmulRir() : buf(kRirSmemFloats, 0.f) {
        float* p = buf.data();
        s.a_re = p; p += swz_len(kRirH); s.a_im = p; p += swz_len(kRirH);
        s.b_re = p; p += swz_len(kRirH); s.b_im = p;
    }
};
struct SrcX {
    const float* x; long long off, L; float scale;
    float operator()(int n) const { long long i = off + n; return (i >= 0 && i < L) ? x[i] * scale : 0.f; }
};
}  // namespace

extern "C" {
void emul_rir_spectrum(const float* ir, int K, const float* tw4096, const float* w8192, float* spec) {
    EmulRir e;
    const cf* tw = reinterpret_cast<const cf*>(tw4096);
    const cf* w = reinterpret_cast<const cf*>(w8192);
    SrcX src{ir, 0, K, 1.f};
    RirStore st{nullptr, 0, 0, 0.f};
    for (int ph = 0; ph < 4; ++ph)
        for (int tid = 0; tid < kRirThreads; ++tid) rir_block_phase<false>(ph, tid, tw, w, nullptr, e.s, src, st);
    for (int tid = 0; tid < kRirThreads; ++tid)
        rir_unpack_spectrum(tid, SwzLoad{e.s.b_re, e.s.b_im}, w, reinterpret_cast<cf*>(spec));
}

void emul_rir_correlate(const float* x, long long L, const float* spec, int K, const float* tw4096,
                        const float* w8192, float* y) {
    EmulRir e;
    RirGeom g = rir_geom(L, K);
    const cf* tw = reinterpret_cast<const cf*>(tw4096);
    const cf* w = reinterpret_cast<const cf*>(w8192);
    const long long nblk = (g.nout + g.valid - 1) / g.valid;
    for (long long b = 0; b < nblk; ++b) {
        long long i0 = b * g.valid;
        SrcX src{x, i0 - g.pad, L, 1.f};
        RirStore st{y + i0, 0, (int)std::min<long long>(g.valid, g.nout - i0), 1.0f / kRirN};
        for (int ph = 0; ph < kRirPhases; ++ph)
            for (int tid = 0; tid < kRirThreads; ++tid)
                rir_block_phase<true>(ph, tid, tw, w, reinterpret_cast<const cf*>(spec), e.s, src, st);
    }
}

void emul_rir_adjoint(const float* ybar, long long L, const float* spec, int K, const float* tw4096,
                      const float* w8192, float scale, float* xbar) {
    EmulRir e;
    RirGeom g = rir_geom(L, K);
    const cf* tw = reinterpret_cast<const cf*>(tw4096);
    const cf* w = reinterpret_cast<const cf*>(w8192);
    const long long nblk = (L + g.valid - 1) / g.valid;
    for (long long b = 0; b < nblk; ++b) {
        long long j0 = b * g.valid;
        SrcX src{ybar, j0 + g.pad - (K - 1), g.nout, scale};
        RirStore st{xbar + j0, K - 1, (int)std::min<long long>(g.valid, L - j0), 1.0f / kRirN};
        for (int ph = 0; ph < kRirPhases; ++ph)
            for (int tid = 0; tid < kRirThreads; ++tid)
                rir_block_phase<false>(ph, tid, tw, w, reinterpret_cast<const cf*>(spec), e.s, src, st);
    }
}
}

// ------------------------------------------------------------------------------------------------------------------
// Polyphase sinc-resampling FIR and its adjoint (diffmusic_b200/csrc/fir_poly.cuh), same chunk loops as guidance_ops.cu
#include "../../diffmusic_b200/csrc/fir_poly.cuh"

extern "C" {
void emul_resample_fwd(const float* x, long long L, const float* kernel, int taps, int orig, int width, float* y,
                       long long Ly) {
    const int CH = 2048;
    const int span = orig * (CH + kFirR) + taps;
    std::vector<float> xs(fir_padded_len(span)), outs(CH + kFirR);
    for (long long o0 = 0; o0 < Ly; o0 += CH) {
        const int no = (int)std::min<long long>(CH, Ly - o0);
        const long long x_lo = (long long)orig * o0 - width;
        for (int i = 0; i < span; ++i) { long long g = x_lo + i; xs[fir_pad(i)] = (g >= 0 && g < L) ? x[g] : 0.f; }
        for (int j0 = 0; j0 < no; j0 += kFirR) {
            float acc[kFirR];
            if (orig == 2 && taps == 28) fir_fwd4<2, 28>(xs.data(), kernel, taps, orig, j0, acc);
            else if (orig == 10 && taps == 132) fir_fwd4<10, 132>(xs.data(), kernel, taps, orig, j0, acc);
            else fir_fwd4(xs.data(), kernel, taps, orig, j0, acc);
            for (int c = 0; c < kFirR; ++c) outs[j0 + c] = acc[c];
        }
        for (int t = 0; t < no; ++t) y[o0 + t] = outs[t];
    }
}

void emul_resample_adjoint(const float* ybar, long long Ly, const float* kernel, int taps, int orig, int width,
                           float scale, float* xbar, long long L) {
    const int unit = orig * kFirR;
    const int chunk = unit * ((2048 + unit - 1) / unit);
    const int span = (taps + chunk) / orig + kFirR + 2;
    std::vector<float> ybuf(span + 1, 0.f), outs(chunk);
    float* ys = ybuf.data() + 1;  // ys[-1] = 0
    for (long long i0 = 0; i0 < L; i0 += chunk) {
        const int ni = (int)std::min<long long>(chunk, L - i0);
        const long long num = i0 + width - taps + 1;
        const long long j_base = ceil_div_ll(num, orig);
        const int A = (int)(num - j_base * orig);
        for (int i = 0; i < span; ++i) { long long o = j_base + i; ys[i] = (o >= 0 && o < Ly) ? ybar[o] * scale : 0.f; }
        for (int wi = 0; wi < chunk / kFirR; ++wi) {
            const int phi = wi % orig, u = wi / orig, t0 = phi + orig * kFirR * u;
            float acc[kFirR];
            if (orig == 2 && taps == 28) fir_adj4<2, 28>(ys, kernel, taps, orig, A, t0, acc);
            else if (orig == 10 && taps == 132) fir_adj4<10, 132>(ys, kernel, taps, orig, A, t0, acc);
            else fir_adj4(ys, kernel, taps, orig, A, t0, acc);
            for (int c = 0; c < kFirR; ++c) outs[t0 + orig * c] = acc[c];
        }
        for (int t = 0; t < ni; ++t) xbar[i0 + t] = outs[t];
    }
}
}

// register-window scale-2 bodies (fir2_fwd8 / fir2_adj8), windows built exactly as the kernels build them
extern "C" {
void emul_resample2_fwd(const float* x, long long L, const float* kernel, float* y, long long Ly) {
    float h[kFir2Taps];
    for (int k = 0; k < kFir2Taps; ++k) h[k] = kernel[k];
    for (long long j0 = 0; j0 < Ly; j0 += kFir2Out) {
        float win[kFir2FwdWin], out[kFir2Out];
        for (int n = 0; n < kFir2FwdWin; ++n) { long long g = 2 * j0 - 16 + n; win[n] = (g >= 0 && g < L) ? x[g] : 0.f; }
        fir2_fwd8(win, h, out);
        for (int c = 0; c < kFir2Out; ++c) if (j0 + c < Ly) y[j0 + c] = out[c];
    }
}
void emul_resample2_adjoint(const float* ybar, long long Ly, const float* kernel, float scale, float* xbar, long long L) {
    float h[kFir2Taps];
    for (int k = 0; k < kFir2Taps; ++k) h[k] = kernel[k];
    for (long long i0 = 0; i0 < L; i0 += kFir2Out) {
        float win[kFir2AdjWin], out[kFir2Out];
        for (int n = 0; n < kFir2AdjWin; ++n) { long long j = i0 / 2 - 8 + n; win[n] = (j >= 0 && j < Ly) ? ybar[j] : 0.f; }
        fir2_adj8(win, h, out);
        for (int c = 0; c < kFir2Out; ++c) if (i0 + c < L) xbar[i0 + c] = out[c] * scale;
    }
}
}

// persistent forward kernel: stage a chunk's input span cell by cell into the swizzled buffer, then every thread reads
// its window through rs2_win_cell -- the data path of resample2_fwd_stream_kernel (guidance_ops.cu)
extern "C" {
void emul_resample2_fwd_stream(const float* x, long long L, const float* kernel, float* y, long long Ly) {
    float h[kFir2Taps];
    for (int k = 0; k < kFir2Taps; ++k) h[k] = kernel[k];
    std::vector<float> buf(kRs2BufFloats);
    const long long chunks = (Ly + kRs2ChunkOut - 1) / kRs2ChunkOut;
    for (long long ch = 0; ch < chunks; ++ch) {
        const long long j0c = ch * kRs2ChunkOut, x0 = 2 * j0c - 16;
        std::fill(buf.begin(), buf.end(), -1e30f);  // poison: every cell a window reads must have been staged
        for (int c = 0; c < kRs2Cells; ++c)
            for (int e = 0; e < 4; ++e) {
                const long long g = x0 + 4 * c + e;
                buf[4 * rs2_cell(c) + e] = (g >= 0 && g < L) ? x[g] : 0.f;
            }
        for (int t = 0; t < kRs2Threads; ++t) {
            const long long j0 = j0c + (long long)t * kFir2Out;
            if (j0 >= Ly) continue;
            float win[kFir2FwdWin], out[kFir2Out];
            for (int q = 0; q < kFir2FwdWin / 4; ++q)
                for (int e = 0; e < 4; ++e) win[4 * q + e] = buf[4 * rs2_win_cell(t, q) + e];
            fir2_fwd8(win, h, out);
            for (int c = 0; c < kFir2Out; ++c)
                if (j0 + c < Ly) y[j0 + c] = out[c];
        }
    }
}
// bank-conflict audit: per quarter-warp (8 lanes) the 16-byte cells of one 128-bit access must fall into 8 different
// bank groups (cell index mod 8).  Returns the number of conflicting accesses.
int emul_rs2_audit() {
    int bad = 0;
    for (int c = 0; c < kRs2Cells; ++c) bad += rs2_cell(c) < 0 || rs2_cell(c) >= kRs2BufFloats / 4;
    for (int t = 0; t < kRs2Threads; ++t)
        for (int q = 0; q < 12; ++q) bad += rs2_win_cell(t, q) != rs2_cell(4 * t + q);
    for (int c0 = 0; c0 + 8 <= kRs2Cells + 7; c0 += 8) {  // staging: 8 consecutive cells per quarter-warp
        int seen = 0;
        for (int i = 0; i < 8; ++i) seen |= 1 << (rs2_cell(c0 + i) & 7);
        bad += seen != 0xff;
    }
    for (int t0 = 0; t0 < kRs2Threads; t0 += 8)  // window reads: lanes t0 .. t0 + 7, cell q each
        for (int q = 0; q < 12; ++q) {
            int seen = 0;
            for (int i = 0; i < 8; ++i) seen |= 1 << (rs2_win_cell(t0 + i, q) & 7);
            bad += seen != 0xff;
        }
    return bad;
}
}

// exhaustive check of the closed-form swizzled addresses (fft_core.cuh) against swz() of the logical index
extern "C" int emul_check_swizzle_forms() {
    int bad = 0;
    for (int j = 0; j < 64; ++j)
        for (int r = 0; r < 8; ++r) {
            bad += ld_addr<512>(swz(j), r) != swz(j + 64 * r);
            bad += st_addr<512, 1>(st_base<512, 1>(j), r) != swz(8 * j + r);
            bad += st_addr<512, 8>(st_base<512, 8>(j), r) != swz((j / 8) * 64 + j % 8 + 8 * r);
            bad += st_addr<512, 64>(st_base<512, 64>(j), r) != swz(j + 64 * r);
        }
    for (int j = 0; j < 512; ++j)
        for (int r = 0; r < 8; ++r) {
            bad += ld_addr<4096>(swz(j), r) != swz(j + 512 * r);
            bad += st_addr<4096, 1>(st_base<4096, 1>(j), r) != swz(8 * j + r);
            bad += st_addr<4096, 8>(st_base<4096, 8>(j), r) != swz((j / 8) * 64 + j % 8 + 8 * r);
            bad += st_addr<4096, 64>(st_base<4096, 64>(j), r) != swz((j / 64) * 512 + j % 64 + 64 * r);
            bad += st_addr<4096, 512>(st_base<4096, 512>(j), r) != swz(j + 512 * r);
        }
    return bad;
}
