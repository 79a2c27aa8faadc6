// include/digital_filter.hpp -- the reference's C++ surface over the B200 library.
//
// Drop-in for digital-filtering-c++/df/df.hpp: the same names a caller uses
//     DFConfig config; DIGITAL_FILTER df(config); df.filter(dt);  df.u.fluc[j*Nz + k] ...
// (test/cpp-main.cpp:6-17), bound to the C ABI in dfb200.h.  Header-only; link with -ldfb200.
//
//   struct DFConfig          df.hpp:38-49   (field for field; `dc_config` = README.md:21's spelling)
//   struct FilterField       df.hpp:24-34   (the members a caller reads: fluc, filt, N_ys, N_zs, Ny_max, Nz_max, Lt)
//   class  DIGITAL_FILTER    df.hpp:52-125  (public u, v, w; DIGITAL_FILTER(DFConfig); filter(double))
//
// Differences a caller can observe, all deliberate:
//   * filter() neither prints "Filtering took ..." nor rewrites ../files/cpp_vel_fluc.csv every call
//     (df.cpp:462-467);
//   * T_fluc / rho_fluc are private in the reference (df.hpp:59): read them through T_fluc() / rho_fluc();
//   * errors throw std::runtime_error with the library's message instead of printing to cerr and
//     carrying on half-initialised (df.cpp:225-228, 492-495);
//   * the per-stage public methods (generate_white_noise, filtering_sweeps, correlate_fields,
//     apply_RST_scaling, get_rho_T_fluc; df.hpp:96-101) do not exist: the stages are fused on the GPU.
//     get_rms() / plot_rms() / write_csv() / write_tecplot() (df.hpp:108-118) do: the reference's own driver calls get_rms()
//     (test/cpp-main.cpp:17), and that file compiles unchanged against include/df/df.hpp.
//   * the host vectors are page-locked (dfb_host_register) so that the five copies per filter() run at the PCIe rate;
//   * the noise is the counter-based pcg32 stream of include/dfb_rng_spec.h (seedable, reproducible)
//     instead of a random_device-seeded process-wide static (df.cpp:334).
#pragma once
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>
#include "dfb200.h"

typedef std::vector<double> Vector;

struct FilterField {
    Vector filt, fluc;                 // Ny*Nz, row-major j*Nz + k  (df.hpp:28)
    Vector rms_added, rms;             // df.hpp:27: filled by get_rms() (sums of squares on the device, df.cpp:571-582)
    std::vector<int> N_ys, N_zs;       // df.hpp:30
    double Lt = 0;                     // df.hpp:32
    int Nz_max = 0, Ny_max = 0;        // df.hpp:33
};

struct DFConfig {                      // df.hpp:38-49
    double d_i, rho_e, U_e, mu_e;
    int vel_file_offset, vel_file_N_values;
    std::string grid_file, vel_fluc_file;
};
typedef DFConfig dc_config;

// Everything the reference hard-codes (df.cpp:7-16, 73-74, 224) made configurable; defaults = the reference.
struct DFConfigEx {
    DFConfig base{0, 0, 0, 0, 0, 0, "", ""};
    bool honor_flow_config = false;    // take d_i, rho_e, U_e, mu_e from `base` (the Fortran behaviour, df.f90:80-87)
    std::string line_file;             // "" -> "../line.dat" (df.cpp:16)
    int Ny = 0, Nz = 0;                // 0,0 -> the reference's made-up grid (df.cpp:73-74)
    bool geom_per_row = true;
    const double *yc = nullptr, *dy = nullptr, *dz = nullptr, *rows = nullptr, *scales = nullptr;
    const int *N_y = nullptr, *N_z = nullptr;
    std::uint64_t seed = 0;
    int noise_mode = DFB_NOISE_GENERATE;
    int device = -1, plane_id = 0, k_begin = 0, k_end = 0;
    bool fetch_every_step = true;      // copy u', v', w', T', rho' into the host vectors after each filter()
};

class DIGITAL_FILTER {
  public:
    FilterField u, v, w;               // df.hpp:87

    // df.hpp:89 -- like the reference (df.cpp:7-16) this constructor IGNORES `config` and builds the
    // hard-coded default plane from ../files/RST.dat and ../line.dat.
    explicit DIGITAL_FILTER(DFConfig) { init(DFConfigEx()); }
    explicit DIGITAL_FILTER(const DFConfigEx& cfg) { init(cfg); }
    DIGITAL_FILTER(const DIGITAL_FILTER&) = delete;
    DIGITAL_FILTER& operator=(const DIGITAL_FILTER&) = delete;
    ~DIGITAL_FILTER() {
        for (void* p : pinned_) dfb_host_unregister(p);
        if (h_) dfb_destroy(h_);
    }

    // df.hpp:100, df.cpp:449-468
    void filter(double dt_input) {
        dt = dt_input;
        if (fetch_)
            check(dfb_filter_to_host(h_, dt, u.fluc.data(), v.fluc.data(), w.fluc.data(), T_fluc_.data(), rho_fluc_.data()));
        else
            check(dfb_filter(h_, dt));
    }
    // explicit fetch for fetch_every_step = false (also refreshes .filt)
    void fetch() {
        FilterField* F[3] = {&u, &v, &w};
        for (int f = 0; f < 3; ++f) {
            check(dfb_get_field(h_, DFB_U_FLUC + f, F[f]->fluc.data(), 0));
            check(dfb_get_field(h_, DFB_U_FILT + f, F[f]->filt.data(), 0));
        }
        check(dfb_get_field(h_, DFB_T_FLUC, T_fluc_.data(), 0));
        check(dfb_get_field(h_, DFB_RHO_FLUC, rho_fluc_.data(), 0));
    }
    // df.hpp:112, df.cpp:584-611 -- the reference's validation driver: 500 steps of dt = 1e-5 accumulating the squares of the five
    // fields (here: on the device, in the z-sweep's epilogue), then plot_rms().  This is the call test/cpp-main.cpp:17 makes.
    void get_rms() {
        dt = 1e-5;
        check(dfb_get_rms(h_, 500, dt));
        FilterField* F[3] = {&u, &v, &w};
        std::int64_t cnt = 0;
        for (int f = 0; f < 3; ++f) {
            F[f]->rms_added.assign(n_cells, 0.0); F[f]->rms.assign(n_cells, 0.0);
            check(dfb_stats_get(h_, f, 0, F[f]->rms_added.data(), &cnt));
            check(dfb_stats_get(h_, f, 1, F[f]->rms.data(), &cnt));
        }
        T_rms_.assign(n_cells, 0.0); rho_rms_.assign(n_cells, 0.0);
        check(dfb_stats_get(h_, 3, 1, T_rms_.data(), &cnt));
        check(dfb_stats_get(h_, 4, 1, rho_rms_.data(), &cnt));
        rms_counter = (int)cnt;
        fetch();                       // the fields of the last step, as the reference leaves them
        plot_rms();
    }
    // df.hpp:113, df.cpp:613-675: writes ../files/cpp_vel_fluc_rms.csv and says so
    void plot_rms() {
        const std::string filename = "../files/cpp_vel_fluc_rms.csv";
        check(dfb_write_rms_csv(h_, filename.c_str()));
        std::printf("Finished plotting to file: %s\n", filename.c_str());
    }
    void write_csv(const std::string& filename) { check(dfb_write_csv(h_, filename.c_str())); std::printf("CSV written to %s\n", filename.c_str()); }   // df.cpp:764-803
    void write_tecplot(const std::string& filename) { check(dfb_write_tecplot(h_, filename.c_str())); std::printf("Finished plotting.\n"); }           // df.cpp:712-762
    const Vector& T_rms() const { return T_rms_; }
    const Vector& rho_rms() const { return rho_rms_; }
    int rms_counter = 0;               // df.hpp:63

    const Vector& T_fluc() const { return T_fluc_; }
    const Vector& rho_fluc() const { return rho_fluc_; }
    int get_Ny() const { return Ny; }
    int get_Nz() const { return Nz; }
    dfb_handle handle() const { return h_; }
    double dt = 0;                     // df.hpp:64

  private:
    int Ny = 0, Nz = 0, n_cells = 0;   // df.hpp:56
    Vector rho_fluc_, T_fluc_;         // df.hpp:59
    Vector T_rms_, rho_rms_;           // df.hpp:84
    std::vector<void*> pinned_;
    dfb_handle h_ = nullptr;
    bool fetch_ = true;

    static void check(int rc) {
        if (rc != DFB_OK) throw std::runtime_error(std::string("DIGITAL_FILTER: ") + dfb_last_error());
    }
    void init(const DFConfigEx& c) {
        dfb_config k;
        dfb_config_init(&k);
        k.d_i = c.base.d_i; k.rho_e = c.base.rho_e; k.U_e = c.base.U_e; k.mu_e = c.base.mu_e;
        k.vel_file_offset = c.base.vel_file_offset; k.vel_file_N_values = c.base.vel_file_N_values;
        k.grid_file = c.base.grid_file.empty() ? nullptr : c.base.grid_file.c_str();
        k.vel_fluc_file = c.base.vel_fluc_file.empty() ? nullptr : c.base.vel_fluc_file.c_str();
        k.line_file = c.line_file.empty() ? nullptr : c.line_file.c_str();
        k.honor_flow_config = c.honor_flow_config;
        k.Ny = c.Ny; k.Nz = c.Nz; k.geom_per_row = c.geom_per_row;
        k.yc = c.yc; k.dy = c.dy; k.dz = c.dz; k.rows = c.rows; k.scales = c.scales; k.N_y = c.N_y; k.N_z = c.N_z;
        k.seed = c.seed; k.noise_mode = c.noise_mode; k.device = c.device; k.plane_id = c.plane_id;
        k.k_begin = c.k_begin; k.k_end = c.k_end;
        fetch_ = c.fetch_every_step;
        check(dfb_create(&k, &h_));
        check(dfb_dims(h_, &Ny, &Nz));
        n_cells = Ny * Nz;
        rho_fluc_.assign(n_cells, 0.0); T_fluc_.assign(n_cells, 0.0);
        FilterField* F[3] = {&u, &v, &w};
        double Lt[3];
        check(dfb_get_table(h_, 10, 0, Lt, 3));
        for (int f = 0; f < 3; ++f) {
            F[f]->filt.assign(n_cells, 0.0); F[f]->fluc.assign(n_cells, 0.0);
            F[f]->N_ys.assign(n_cells, 0); F[f]->N_zs.assign(n_cells, 0);
            check(dfb_get_half_widths(h_, f, 0, F[f]->N_ys.data()));
            check(dfb_get_half_widths(h_, f, 1, F[f]->N_zs.data()));
            std::int64_t v64 = 0;
            check(dfb_info(h_, 0, f, &v64)); F[f]->Ny_max = (int)v64;
            check(dfb_info(h_, 1, f, &v64)); F[f]->Nz_max = (int)v64;
            F[f]->Lt = Lt[f];
        }
        // page-lock what filter() copies into every step (a failure only costs speed: pageable copies still work)
        Vector* out[5] = {&u.fluc, &v.fluc, &w.fluc, &T_fluc_, &rho_fluc_};
        for (Vector* a : out)
            if (!a->empty() && dfb_host_register(a->data(), a->size() * sizeof(double)) == DFB_OK) pinned_.push_back(a->data());
        if (c.noise_mode == DFB_NOISE_GENERATE) fetch();   // the constructor's first step (df.cpp:57-62)
    }
};
