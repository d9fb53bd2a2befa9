// heat_3d_slabs -- examples/heat/heat_3d of the reference on several GPUs (or several virtual ranks of one GPU)
// from one C++17 process: the shipped initial state is projected and solved by the single-GPU class, then the
// explicit ADS steps run on z-slabs (ads::slab_cluster).  Prints the same line as heat_3d.
//     heat_3d_slabs [elements] [steps] [ranks] [p]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ads/slabs.hpp"
#include "heat_3d.hpp"

namespace {

// heat_3d with its set-up phase exposed: before() gives the shipped initial state
struct heat_3d_setup : ads::problems::heat_3d {
    using heat_3d::heat_3d;
    ads::dimension& dim(int a) { return a == 0 ? x : a == 1 ? y : z; }
};

}  // namespace

int main(int argc, char* argv[]) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 12;
    const int nsteps = argc > 2 ? std::atoi(argv[2]) : 100;
    const int ranks = argc > 3 ? std::atoi(argv[3]) : 2;
    const int p = argc > 4 ? std::atoi(argv[4]) : 2;
    const double dt = 1e-7;
    ads::dim_config dim{p, n};
    ads::config_3d c{dim, dim, dim, ads::timesteps_config{0, dt}, 1};
    heat_3d_setup sim{c};
    sim.run();  // zero steps: before() only -> the shipped initial state
    ads::lin::tensor<double, 3> u = sim.solution();

    const int gpus = adsb_device_count();
    std::vector<int> devices(ranks, 0);
    if (gpus >= ranks)
        for (int r = 0; r < ranks; ++r) devices[r] = r;
    ads::slab_cluster cluster{sim.dim(0), sim.dim(1), sim.dim(2), devices};
    adsb_substep sub{};
    sub.form = ads::make_form(1.0, {dt, dt, dt});
    sub.fix_axis = sub.fix_buf = -1;
    cluster.commit({sub});
    cluster.set_state(u);
    cluster.advance(nsteps);
    cluster.get_state(u);

    double sum = 0, sq = 0;
    for (int i = 0; i < u.size(); ++i) {
        sum += u.data()[i];
        sq += u.data()[i] * u.data()[i];
    }
    std::printf("heat_3d p=%d n=%d steps=%d: sum(u) = %.14f  |u|_2 = %.14f\n", p, n, nsteps, sum, std::sqrt(sq));
    std::printf("ranks = %d (%s), slab bounds:", ranks, cluster.virtual_ranks() ? "virtual, one device" : "one device each");
    for (int b : cluster.bounds()) std::printf(" %d", b);
    std::printf("\n");
}
