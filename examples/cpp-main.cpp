// examples/cpp-main.cpp -- the reference's own usage example (digital-filtering-c++/test/cpp-main.cpp:3-19)
// compiled against the B200 facade instead of df/df.hpp: same three lines of caller code.
// Run from a directory where ../files/RST.dat and ../line.dat resolve, like the reference (df.cpp:16,224).
//   g++ -std=c++17 -Iinclude examples/cpp-main.cpp -Ldigital-filtering_b200/lib -ldfb200 -o cpp-test
#include <cmath>
#include <cstdio>
#include "digital_filter.hpp"

int main() {
    // Create configuration struct
    DFConfig config;

    // Constructor
    DIGITAL_FILTER df(config);

    // Call filter procedure with timestep (BASELINE.json configs[0]: 100 filter(dt) steps)
    double dt = 1e-5;
    double acc = 0.0;
    for (int i = 0; i < 100; ++i) {
        df.filter(dt);
        for (double x : df.u.fluc) acc += x * x;
    }
    std::printf("Ny=%d Nz=%d  rms(u') over 100 steps = %.6f m/s\n", df.get_Ny(), df.get_Nz(),
                std::sqrt(acc / (100.0 * df.u.fluc.size())));
    return 0;
}
