// ads/simulation/simulation_base.hpp -- the time loop (include/ads/simulation/simulation_base.hpp:16-24,
// src/ads/simulation/simulation_base.cpp:11-20).
#ifndef ADSB_ADS_SIMULATION_BASE_HPP
#define ADSB_ADS_SIMULATION_BASE_HPP

#include "ads/simulation/config.hpp"

namespace ads {

class simulation_base {
protected:
    timesteps_config steps;

    virtual void before() { }
    virtual void after() { }
    virtual void before_step(int /*iter*/, double /*t*/) { }
    virtual void step(int /*iter*/, double /*t*/) { }
    virtual void after_step(int /*iter*/, double /*t*/) { }

public:
    explicit simulation_base(const timesteps_config& steps) : steps{steps} { }
    virtual ~simulation_base() = default;

    void run() {
        before();
        for (int i = 0; i < steps.step_count; ++i) {
            const double t = i * steps.dt;
            before_step(i, t);
            step(i, t);
            after_step(i, t);
        }
        after();
    }
};

}  // namespace ads

#endif
