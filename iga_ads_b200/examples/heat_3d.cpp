// examples/heat/heat_3d.cpp of the reference, with optional arguments: elements, steps, method.
#include "heat_3d.hpp"

#include <cstdio>
#include <cstdlib>

int main(int argc, char* argv[]) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 12;
    const int nsteps = argc > 2 ? std::atoi(argv[2]) : 100;
    const int method = argc > 3 ? std::atoi(argv[3]) : ADSB_RHS_COLLAPSED;
    ads::dim_config dim{2, n};
    ads::timesteps_config steps{nsteps, 1e-7};
    int ders = 1;

    ads::config_3d c{dim, dim, dim, steps, ders};
    ads::problems::heat_3d sim{c, method};
    sim.run();

    const auto& u = sim.solution();
    double sum = 0, sq = 0;
    for (int i = 0; i < u.size(); ++i) {
        sum += u.data()[i];
        sq += u.data()[i] * u.data()[i];
    }
    std::printf("heat_3d p=2 n=%d steps=%d: sum(u) = %.14f  |u|_2 = %.14f\n", n, nsteps, sum, std::sqrt(sq));
}
