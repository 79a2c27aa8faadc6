// examples/facade_bench.cpp -- end-to-end rate of the reference-facing C++ call, measured the way a C++ caller sees it:
// DIGITAL_FILTER (include/digital_filter.hpp) on a caller-supplied plane, K x df.filter(dt) with the five fields landing in the
// facade's own std::vector members every step (page-locked by the facade).  bench.py writes the plane file, builds and runs this.
//   usage: facade_bench <plane.bin> <steps> <warmup> [device]     prints one line: FACADE cells steps seconds checksum
//   plane.bin: int32 Ny, Nz; then doubles yc[Ny], dy[Ny], dz[Ny], rows[8*Ny], scales[9]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "digital_filter.hpp"

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: %s plane.bin steps warmup [device]\n", argv[0]); return 2; }
    std::FILE* fp = std::fopen(argv[1], "rb");
    if (!fp) { std::perror(argv[1]); return 3; }
    int dims[2];
    if (std::fread(dims, sizeof(int), 2, fp) != 2) return 4;
    const int Ny = dims[0], Nz = dims[1];
    std::vector<double> yc(Ny), dy(Ny), dz(Ny), rows(8 * (size_t)Ny), scales(9);
    auto rd = [&](std::vector<double>& v) { return std::fread(v.data(), sizeof(double), v.size(), fp) == v.size(); };
    if (!rd(yc) || !rd(dy) || !rd(dz) || !rd(rows) || !rd(scales)) return 5;
    std::fclose(fp);
    const int steps = std::atoi(argv[2]), warmup = std::atoi(argv[3]);

    DFConfigEx cfg;
    cfg.base.d_i = 0.0013; cfg.base.U_e = 869.1;
    cfg.honor_flow_config = true;
    cfg.Ny = Ny; cfg.Nz = Nz; cfg.geom_per_row = true;
    cfg.yc = yc.data(); cfg.dy = dy.data(); cfg.dz = dz.data(); cfg.rows = rows.data(); cfg.scales = scales.data();
    cfg.seed = 20261018;
    cfg.device = argc > 4 ? std::atoi(argv[4]) : -1;
    DIGITAL_FILTER df(cfg);
    const double dt = 1e-7;
    for (int i = 0; i < warmup; ++i) df.filter(dt);
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < steps; ++i) df.filter(dt);            // returns with u.fluc ... rho_fluc() of this step in host memory
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double chk = 0.0;
    for (size_t i = 0; i < df.u.fluc.size(); i += 4097) chk += df.u.fluc[i] + df.rho_fluc()[i];
    std::printf("FACADE %lld %d %.9f %.17g\n", (long long)df.u.fluc.size(), steps, secs, chk);
    return 0;
}
