// Multi-GPU from C++ (one process per GPU, launched by torchrun): the cfg5 pipeline of BASELINE.json --
// m = mean<float>(a, {0}), v = variance<float>(a, {0}) (two-pass), out = exp(a - m) -- on a row-sharded matrix
// through xtb::dist, checked against the single-process evaluation of the FULL matrix on the same GPU and
// against the real xtensor on the host.  Also: allreduce of sum / amax / prod reducers over the sharded axis,
// xt::initial applied once, a reduction over a non-sharded axis (no exchange), ragged row split.
// Model: reducers/xblockwise_reducer.hpp:154-185 (partial -> merge -> finalize).
//   torchrun --nproc-per-node 2 --no-python tests/cpp/_build/test_dist      (or: RANK/WORLD_SIZE/... set by hand)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

#include <xtb200/xtensor_b200.hpp>
#include <xtensor/generators/xbuilder.hpp>

static int g_failed = 0;
#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++g_failed; } \
    } while (0)

template <class A, class B> bool same_bits(const A& a, const B& b)
{
    if (a.size() != b.size()) return false;
    return std::memcmp(a.data(), b.data(), a.size() * sizeof(typename A::value_type)) == 0;
}
template <class A, class B> double max_rel(const A& a, const B& b)
{
    double m = 0;
    auto ib = b.begin();
    for (auto ia = a.begin(); ia != a.end(); ++ia, ++ib)
        m = std::max(m, std::fabs(double(*ia) - double(*ib)) / std::max(1e-300, std::fabs(double(*ib))));
    return m;
}

int main()
{
    xtb::dist::communicator cm;
    try { cm = xtb::dist::init_from_env(); }
    catch (std::exception& e) { std::printf("init failed: %s\n", e.what()); return 2; }
    const int rank = cm.rank, world = cm.world;

    const std::size_t rows = 1031, cols = 768;          // ragged split over the ranks
    // integer-valued data: sums are exact in any order, so sharded == single == reference bit for bit
    std::mt19937_64 gen(9);
    std::uniform_int_distribution<int> dist(-8, 8);
    xt::xtensor<float, 2> a = xt::zeros<float>({rows, cols});
    for (auto& x : a) x = float(dist(gen));
    const auto [r0, r1] = xtb::dist::row_block(rows, rank, world);
    xt::xtensor<float, 2> a_local = xt::view(a, xt::range(r0, r1), xt::all());
    xtb::xtensor<float, 2> da = xtb::to_device(a_local), dfull = xtb::to_device(a);
    const std::vector<std::size_t> ax0 = {0};

    // sums / extremes over the sharded axis: one fused local-reduce + exchange per call
    {
        auto s = xtb::dist::allreduce(xt::sum(da, {0}));
        xt::xtensor<float, 1> hs = xt::sum(a, {0});
        CHECK(same_bits(xtb::to_host(s), hs));
        auto mx = xtb::dist::allreduce(xt::amax(da, {0}));
        xt::xtensor<float, 1> hmx = xt::amax(a, {0});
        CHECK(same_bits(xtb::to_host(mx), hmx));
        auto s01 = xtb::dist::allreduce(xt::sum(da, {0, 1}));
        CHECK(xtb::to_host(s01)() == xt::sum(a, {0, 1})());
        // xt::initial is merged ONCE, after the exchange
        auto si = xtb::dist::allreduce(xt::sum(da, {0}, xt::initial(100.0f)));
        xt::xtensor<float, 1> hsi = xt::sum(a, {0}, xt::initial(100.0f));
        CHECK(same_bits(xtb::to_host(si), hsi));
        // non-sharded axis: no exchange, the result stays sharded by rows
        xtb::xtensor<float, 1> s1 = xt::sum(da, {1});
        xt::xtensor<float, 1> hs1 = xt::sum(a_local, {1});
        CHECK(same_bits(xtb::to_host(s1), hs1));
    }
    // the cfg5 pipeline
    {
        auto m = xtb::dist::mean<float>(da, ax0, rows);
        auto v = xtb::dist::variance<float>(da, ax0, rows);
        xtb::xtensor<float, 2> out;
        xt::noalias(out) = xt::exp(da - m);

        xt::xtensor<float, 1> hm = xt::mean<float>(a, {0});
        xt::xtensor<float, 1> hv = xt::variance<float>(a, {0});
        auto gm = xtb::to_host(m);
        CHECK(same_bits(gm, hm));                                   // exact sums, one division: bit-identical
        // reorder-sensitive: the device must be within 1e-6 of the fp64 value of the same two passes, and at least as
        // close to it as the reference's own sequential fp32 order is (which drifts by ~rows * eps / 2)
        xt::xtensor<double, 2> dev2 = xt::square(xt::cast<double>(a) - xt::cast<double>(hm));
        xt::xtensor<double, 1> v64 = xt::sum(dev2, {0}) / double(rows);
        const double err_dev = max_rel(xtb::to_host(v), v64), err_ref = max_rel(hv, v64);
        CHECK(err_dev <= 1e-6);
        CHECK(err_dev <= std::max(err_ref, 1e-6));
        CHECK(max_rel(xtb::to_host(v), hv) <= std::max(4 * err_ref, 1e-6));
        // single-process evaluation of the full matrix on this GPU
        xtb::xtensor<float, 1> m1 = xt::mean<float>(dfull, {0});
        xtb::xtensor<float, 1> v1 = xt::variance<float>(dfull, {0});
        CHECK(same_bits(gm, xtb::to_host(m1)));
        CHECK(max_rel(xtb::to_host(v), xtb::to_host(v1)) <= 1e-6);
        xtb::xtensor<float, 2> out1;
        xt::noalias(out1) = xt::exp(dfull - m1);
        xt::xtensor<float, 2> ho1 = xtb::to_host(out1);
        xt::xtensor<float, 2> ho1_rows = xt::view(ho1, xt::range(r0, r1), xt::all());
        CHECK(same_bits(xtb::to_host(out), ho1_rows));              // the map of my rows == those rows of the full map
        // default mean (double quotient)
        auto md = xtb::dist::mean(da, ax0, rows);
        xt::xtensor<double, 1> hmd = xt::mean(a, {0});
        CHECK(same_bits(xtb::to_host(md), hmd));
    }
    xtb::sync();
    std::printf("rank %d/%d peer_memory=%d %s\n", rank, world, int(cm.peer_memory), g_failed ? "FAILED" : "OK test_dist");
    xtb::dist::finalize();
    return g_failed ? 1 : 0;
}
