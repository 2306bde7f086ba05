// Host harness for social_navigation_pyenvs_b200/csrc/snp_reset_core.h (the code the CUDA reset kernel runs per environment):
// compiled by tests/test_scenarios.py with g++ and compared with the outputs recorded from the live reference.
//   reset_core_host <scenario> <N> <randomize_attributes> <seed0> <count>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "snp_reset_core.h"

int main(int argc, char **argv) {
    if (argc < 6) return 2;
    snp::ResetParams p;
    p.scenario = atoi(argv[1]); p.N = atoi(argv[2]); p.randomize_attributes = atoi(argv[3]);
    p.circle_radius = 7.0; p.robot_radius = 0.3; p.traffic_length = 14.0; p.traffic_height = 3.0;
    const unsigned seed0 = (unsigned)atoll(argv[4]);
    const int count = atoi(argv[5]);
    std::vector<uint32_t> state(624);
    std::vector<double> scratch(4 * p.N);
    for (int e = 0; e < count; ++e) {
        snp::Mt19937<snp::SoloGroup> rng{state.data(), 624, 0, snp::SoloGroup{}};
        snp::ResetScratch w{scratch.data(), scratch.data() + p.N, scratch.data() + 2 * p.N, scratch.data() + 3 * p.N};
        std::vector<snp::ResetHuman> out(p.N);
        auto emit = [&](int i, const snp::ResetHuman &h) { out[i] = h; };
        const int scen = snp::reset_generate(p, seed0 + e, rng, w, emit);
        printf("env %d %d %lld\n", e, scen, rng.draws);
        for (int i = 0; i < p.N; ++i)
            printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %d\n", out[i].x, out[i].y, out[i].yaw, out[i].radius, out[i].vd, out[i].g0x,
                   out[i].g0y, out[i].g1x, out[i].g1y, out[i].goal_count);
    }
    return 0;
}
