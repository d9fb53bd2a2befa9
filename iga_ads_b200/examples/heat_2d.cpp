// examples/heat/heat_2d.cpp of the reference, with optional arguments: p, elements, steps.
#include "heat_2d.hpp"

#include <cstdio>
#include <cstdlib>

int main(int argc, char* argv[]) {
    const int p = argc > 1 ? std::atoi(argv[1]) : 2;
    const int n = argc > 2 ? std::atoi(argv[2]) : 40;
    const int nsteps = argc > 3 ? std::atoi(argv[3]) : 100;
    ads::dim_config dim{p, n};
    ads::timesteps_config steps{nsteps, 1e-5};
    ads::config_2d c{dim, dim, steps, 1};
    ads::problems::heat_2d sim{c};
    sim.run();
    const auto& u = sim.solution();
    double sum = 0;
    for (int i = 0; i < u.size(); ++i) sum += u.data()[i];
    std::printf("heat_2d p=%d n=%d steps=%d: sum(u) = %.14f\n", p, n, nsteps, sum);
}
