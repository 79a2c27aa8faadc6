// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Compiles the reference's own, unmodified translation unit
//   /root/reference/digital-filtering-c++/df/df.cpp
// from where it lies (no source is copied) and wraps the resulting DIGITAL_FILTER object in a
// small extern "C" surface that tests/, __graft_entry__.smoke() and bench.py's CPU legs drive
// through ctypes.  Output: oracle/_ref/libdfref.so (git-ignored, travels to the GPU box).
//
// Two preprocessor shims are needed because HEAD of the reference does not compile as shipped
// (SURVEY.md section 0 / 8c):
//   * df.cpp uses a member `z` (df.cpp:78,100,637,724,786) that df.hpp:67-75 never declares.
//     df.hpp:79 declares `double U_w, rho_w, ...;` and U_w is referenced nowhere else, so the
//     macro below turns that one declaration into `double U_w; Vector z; double U_w_shim_, rho_w...`.
//   * the state we must read/write for noise injection (T_fluc, rho_fluc, dt, Ny, Nz, R**, Ms ...)
//     is private (df.hpp:54-84) -> `#define private public` after the std headers are in.
// Nothing else about the reference is altered; every number it produces comes from its own code.

#include <cinttypes>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <cassert>
#include <limits>
#include <iostream>
#include <iterator>
#include <type_traits>
#include <utility>
#include <locale>
#include <new>
#include <stdexcept>
#include <cmath>
#include <vector>
#include <random>
#include <string>
#include <sstream>
#include <iomanip>
#include <fstream>
#include <chrono>
#include <algorithm>
#include <numeric>
#include <initializer_list>
#include <unistd.h>

#ifndef REF_PCG_HPP
#error "REF_PCG_HPP / REF_DF_CPP must be given by oracle/Makefile"
#endif
#include REF_PCG_HPP   // include-guarded: df.hpp's own include of it becomes a no-op

#define private public
#define U_w U_w; Vector z; double U_w_shim_
#include REF_DF_CPP
#undef U_w
#undef private

namespace {

struct Silence {   // the reference prints from its constructor and from filter() (df.cpp:327-329,464)
    std::streambuf* old_out;
    std::ostringstream sink;
    Silence() : old_out(std::cout.rdbuf(sink.rdbuf())) {}
    ~Silence() { std::cout.rdbuf(old_out); }
};

FilterField* field_of(DIGITAL_FILTER* d, int f) {
    return f == 0 ? &d->u : (f == 1 ? &d->v : &d->w);
}

template <class T>
long copy_out(const std::vector<T>& v, T* out, long cap) {
    long n = (long)v.size();
    if (out) {
        if (cap < n) return -n;
        std::memcpy(out, v.data(), sizeof(T) * n);
    }
    return n;
}

}  // namespace

extern "C" {

// Constructs the reference object.  `rundir` must be a directory from which the reference's
// hard-coded relative paths resolve: ../files/RST.dat (df.cpp:224) and ../line.dat (df.cpp:16).
void* ref_create(const char* rundir) {
    if (rundir && chdir(rundir) != 0) return nullptr;
    {
        std::ifstream a("../files/RST.dat"), b("../line.dat");
        if (!a || !b) return nullptr;   // reference would go on half-initialised (df.cpp:225-228)
    }
    Silence s;
    DFConfig cfg;   // ignored by the C++ reference (df.cpp:7-16)
    cfg.d_i = cfg.rho_e = cfg.U_e = cfg.mu_e = 0.0;
    cfg.vel_file_offset = cfg.vel_file_N_values = 0;
    return new DIGITAL_FILTER(cfg);
}

void ref_destroy(void* h) { delete static_cast<DIGITAL_FILTER*>(h); }

void ref_dims(void* h, int* Ny, int* Nz) {
    auto* d = static_cast<DIGITAL_FILTER*>(h);
    *Ny = d->Ny; *Nz = d->Nz;
}

// scalars: 0 d_i, 1 d_v, 2 u_tau, 3 tau_w, 4 dt, 5 U_e, 6 rho_e, 7 mu, 8 N_in
double ref_scalar(void* h, int which) {
    auto* d = static_cast<DIGITAL_FILTER*>(h);
    switch (which) {
        case 0: return d->d_i;  case 1: return d->d_v;  case 2: return d->u_tau;
        case 3: return d->tau_w; case 4: return d->dt;  case 5: return d->U_e;
        case 6: return d->rho_e; case 7: return d->mu;  case 8: return d->N_in;
    }
    return NAN;
}

// per-field scalars: 0 Lt, 1 Iz_inn, 2 Iz_out, 3 Ny_max, 4 Nz_max
double ref_field_scalar(void* h, int f, int which) {
    FilterField* F = field_of(static_cast<DIGITAL_FILTER*>(h), f);
    switch (which) {
        case 0: return F->Lt; case 1: return F->Iz_inn; case 2: return F->Iz_out;
        case 3: return F->Ny_max; case 4: return F->Nz_max;
    }
    return NAN;
}

// row / geometry vectors. which: 0 R11,1 R21,2 R22,3 R33,4 Us,5 Ts,6 rhos,7 Ms,8 Ps,9 yline,
// 10 ydline, 11 yc, 12 dy, 13 dz, 14 y, 15 z, 16 T_fluc, 17 rho_fluc, 18 yin_d, 19 R11_in
long ref_get_vec(void* h, int which, double* out, long cap) {
    auto* d = static_cast<DIGITAL_FILTER*>(h);
    const Vector* v = nullptr;
    switch (which) {
        case 0: v = &d->R11; break;  case 1: v = &d->R21; break;  case 2: v = &d->R22; break;
        case 3: v = &d->R33; break;  case 4: v = &d->Us; break;   case 5: v = &d->Ts; break;
        case 6: v = &d->rhos; break; case 7: v = &d->Ms; break;   case 8: v = &d->Ps; break;
        case 9: v = &d->yline; break; case 10: v = &d->ydline; break; case 11: v = &d->yc; break;
        case 12: v = &d->dy; break;  case 13: v = &d->dz; break;  case 14: v = &d->y; break;
        case 15: v = &d->z; break;   case 16: v = &d->T_fluc; break; case 17: v = &d->rho_fluc; break;
        case 18: v = &d->yin_d; break; case 19: v = &d->R11_in; break;
        default: return 0;
    }
    return copy_out(*v, out, cap);
}

// per-field double vectors. which: 0 by, 1 bz, 2 r_ys, 3 r_zs, 4 filt_old, 5 filt, 6 fluc
static Vector* fvec(FilterField* F, int which) {
    switch (which) {
        case 0: return &F->by; case 1: return &F->bz; case 2: return &F->r_ys; case 3: return &F->r_zs;
        case 4: return &F->filt_old; case 5: return &F->filt; case 6: return &F->fluc;
    }
    return nullptr;
}
long ref_get_fvec(void* h, int f, int which, double* out, long cap) {
    Vector* v = fvec(field_of(static_cast<DIGITAL_FILTER*>(h), f), which);
    return v ? copy_out(*v, out, cap) : 0;
}
long ref_set_fvec(void* h, int f, int which, const double* in, long n) {
    Vector* v = fvec(field_of(static_cast<DIGITAL_FILTER*>(h), f), which);
    if (!v || (long)v->size() != n) return v ? -(long)v->size() : 0;
    std::memcpy(v->data(), in, sizeof(double) * n);
    return n;
}
// per-field int vectors. which: 0 N_ys, 1 N_zs, 2 by_offsets, 3 bz_offsets
long ref_get_ivec(void* h, int f, int which, int* out, long cap) {
    FilterField* F = field_of(static_cast<DIGITAL_FILTER*>(h), f);
    const std::vector<int>* v = which == 0 ? &F->N_ys : which == 1 ? &F->N_zs
                              : which == 2 ? &F->by_offsets : &F->bz_offsets;
    return copy_out(*v, out, cap);
}

// Synthetic shapes: the reference's stage functions consult only members, so
// overwrite the geometry / row tables and re-run ITS allocate_data_structures (df.cpp:120-128)
// and calculate_filter_properties (df.cpp:130-218).  rows = [R11,R21,R22,R33,Us,Ts,rhos,Ms] x Ny.
// scales = per field {Iz_inn, Iz_out, Lt} (9 doubles).
int ref_reshape(void* h, int Ny, int Nz, double d_i, const double* yc, const double* dy,
                const double* dz, const double* rows, const double* scales) {
    auto* d = static_cast<DIGITAL_FILTER*>(h);
    long n = (long)Ny * Nz;
    d->Ny = Ny; d->Nz = Nz; d->n_cells = (int)n; d->d_i = d_i;
    d->yc.assign(yc, yc + n); d->dy.assign(dy, dy + n); d->dz.assign(dz, dz + n);
    d->yc_d.resize(n);
    for (long i = 0; i < n; ++i) d->yc_d[i] = d->yc[i] / d_i;
    d->ydline.resize(Ny); d->yline.resize(Ny);
    for (int j = 0; j < Ny; ++j) { d->ydline[j] = d->yc_d[(long)j * Nz]; d->yline[j] = d->yc[(long)j * Nz]; }
    Vector* tabs[8] = {&d->R11, &d->R21, &d->R22, &d->R33, &d->Us, &d->Ts, &d->rhos, &d->Ms};
    for (int t = 0; t < 8; ++t) tabs[t]->assign(rows + (long)t * Ny, rows + (long)(t + 1) * Ny);
    d->rho_fluc = Vector(n); d->T_fluc = Vector(n);
    d->y.assign((long)(Ny + 1) * (Nz + 1), 0.0);   // only write_csv reads these (df.cpp:779-786)
    d->z.assign((long)(Ny + 1) * (Nz + 1), 0.0);
    for (int f = 0; f < 3; ++f) {
        FilterField* F = field_of(d, f);
        d->allocate_data_structures(*F);
        F->Iz_inn = scales[3 * f + 0]; F->Iz_out = scales[3 * f + 1]; F->Lt = scales[3 * f + 2];
        d->calculate_filter_properties(*F);
    }
    return 0;
}

// The reference's own stages, called exactly in filter()'s order (df.cpp:453-461) but with the
// noise left as the caller injected it (H1 skipped), no print, no CSV.
void ref_step_injected(void* h, double dt) {
    auto* d = static_cast<DIGITAL_FILTER*>(h);
    d->dt = dt;
    for (FilterField* F : {&d->u, &d->v, &d->w}) { d->filtering_sweeps(*F); d->correlate_fields(*F); }
    d->apply_RST_scaling();
    d->get_rho_T_fluc();
}
// The constructor's first step (df.cpp:57-62): sweeps + RST scaling, no blend, no SRA.
void ref_first_step_injected(void* h) {
    auto* d = static_cast<DIGITAL_FILTER*>(h);
    for (FilterField* F : {&d->u, &d->v, &d->w}) d->filtering_sweeps(*F);
    d->apply_RST_scaling();
}
void ref_generate_white_noise(void* h) { static_cast<DIGITAL_FILTER*>(h)->generate_white_noise(); }

// The real thing: DIGITAL_FILTER::filter(dt) (df.cpp:449-468) -- its own noise, its print (silenced),
// its CSV (lands in <rundir>/../files/cpp_vel_fluc.csv).
void ref_filter(void* h, double dt) {
    Silence s;
    static_cast<DIGITAL_FILTER*>(h)->filter(dt);
}

// CPU baseline: n steps of the five stage calls filter() brackets with its stopwatch
// (df.cpp:452-462), excluding the print and the CSV.  Returns seconds; per-stage seconds in
// stage_s[0..4] = noise, sweeps, correlate, RST, SRA (may be null).
double ref_time_steps(void* h, double dt, int nsteps, double* stage_s) {
    auto* d = static_cast<DIGITAL_FILTER*>(h);
    using clk = std::chrono::steady_clock;
    double acc[5] = {0, 0, 0, 0, 0};
    d->dt = dt;
    auto t0 = clk::now();
    for (int s = 0; s < nsteps; ++s) {
        auto a = clk::now();
        d->generate_white_noise();
        auto b = clk::now(); acc[0] += std::chrono::duration<double>(b - a).count();
        for (FilterField* F : {&d->u, &d->v, &d->w}) {
            auto c0 = clk::now();
            d->filtering_sweeps(*F);
            auto c1 = clk::now();
            d->correlate_fields(*F);
            auto c2 = clk::now();
            acc[1] += std::chrono::duration<double>(c1 - c0).count();
            acc[2] += std::chrono::duration<double>(c2 - c1).count();
        }
        auto e = clk::now();
        d->apply_RST_scaling();
        auto f = clk::now(); acc[3] += std::chrono::duration<double>(f - e).count();
        d->get_rho_T_fluc();
        auto g = clk::now(); acc[4] += std::chrono::duration<double>(g - f).count();
    }
    double total = std::chrono::duration<double>(clk::now() - t0).count();
    if (stage_s) for (int i = 0; i < 5; ++i) stage_s[i] = acc[i];
    return total;
}

// ---- the vendored pcg32 itself (pcg_random.hpp:1866), for the RNG gate -------------------------
// has_stream=0 -> pcg32{seed} (the seeding form df.cpp:334 uses); 1 -> pcg32(seed, stream).
// Advances by `delta` (pcg_random.hpp:457-460,640-669) and then draws n outputs.
void ref_pcg32_draw(uint64_t seed, uint64_t stream, int has_stream, uint64_t delta, int n, uint32_t* out) {
    pcg32 rng = has_stream ? pcg32(seed, stream) : pcg32(seed);
    rng.advance(delta);
    for (int i = 0; i < n; ++i) out[i] = rng();
}
// distance between pcg32(seed,stream) after `a` draws and after `b` draws (pcg_random.hpp:671-693)
uint64_t ref_pcg32_distance(uint64_t seed, uint64_t stream, uint64_t a, uint64_t b) {
    pcg32 x(seed, stream), y(seed, stream);
    x.advance(a); y.advance(b);
    return y - x;
}
// libstdc++ std::normal_distribution fed by pcg32{seed}: what generate_white_noise draws (df.cpp:334-339)
void ref_normals(uint64_t seed, int n, double* out) {
    pcg32 rng{seed};
    std::normal_distribution<> dist(0.0, 1.0);
    for (int i = 0; i < n; ++i) out[i] = dist(rng);
}

}  // extern "C"
