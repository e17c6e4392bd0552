// bench_dropin.cpp -- the headline step of bench.py written against the DROP-IN API (the C++ header a user of
// xtensor includes), not the ctypes mirror:
//
//     m   = xt::mean<float>(a, {0})          (xtb::dist::mean_into: partials merged across GPUs, / GLOBAL rows)
//     v   = xt::variance<float>(a, {0})      (two-pass, xtb::dist::variance_into)
//     xt::noalias(out) = xt::exp(a - m)
//
// on this rank's row block of the fp32 (rows, 8192) matrix of BASELINE cfg5.  bench.py starts one copy per rank
// (same RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* environment as torchrun gave it) and takes the end-to-end
// number from here: every e2e step copies the shard from PINNED HOST memory to the device, runs the step, and
// copies out / mean / variance back into pinned host memory.  Rank 0 prints one JSON object.
//
//     bench_dropin <total_rows> <steps> <warmup> <e2e_steps>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>

#include <xtb200/xtensor_b200.hpp>

namespace
{
    constexpr std::size_t COLS = 8192, BLK = 4096;

    double max_over_ranks(double x, int world)
    {
        if (world <= 1) return x;
        xtb::xtensor<double, 1> d;
        d.resize({1});
        xtb::check(xtb_memcpy(d.data(), &x, sizeof(double), XTB_H2D));
        xtb::check(xtb_allreduce(d.data(), 1, XTB_F64, XTB_RED_MAX));
        double r = 0;
        xtb::check(xtb_memcpy(&r, d.data(), sizeof(double), XTB_D2H));
        return r;
    }
}

int main(int argc, char** argv)
{
    const std::size_t total_rows = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 262144;
    const int steps = argc > 2 ? std::atoi(argv[2]) : 20, warmup = argc > 3 ? std::atoi(argv[3]) : 5;
    const int e2e_steps = argc > 4 ? std::atoi(argv[4]) : 3;
    xtb::dist::communicator cm;
    try { cm = xtb::dist::init_from_env(); }
    catch (std::exception& e) { std::printf("{\"error\": \"init: %s\"}\n", e.what()); return 2; }
    const int rank = cm.rank, world = cm.world;
    const std::size_t rows = total_rows / world;
    if (rows % BLK != 0) { std::printf("{\"error\": \"rows per rank must be a multiple of %zu\"}\n", BLK); return 2; }
    try
    {
        // pinned host operands / results: U(-1, 1) block tiled along the rows (the data bench.py times)
        xtb::pinned_xtensor<float, 2> h_a, h_o;
        xtb::pinned_xtensor<float, 1> h_m, h_v;
        h_a.resize({rows, COLS}); h_o.resize({rows, COLS}); h_m.resize({COLS}); h_v.resize({COLS});
        {
            std::mt19937_64 gen(9 + rank);
            std::uniform_real_distribution<float> u(-1.f, 1.f);
            float* p = h_a.data();
            for (std::size_t i = 0; i < BLK * COLS; ++i) p[i] = u(gen);
            for (std::size_t r = BLK; r < rows; r += BLK) std::copy(p, p + BLK * COLS, p + r * COLS);
        }
        xtb::xtensor<float, 2> a, out;
        xtb::xtensor<float, 1> m, v;
        a.resize({rows, COLS}); out.resize({rows, COLS}); m.resize({COLS}); v.resize({COLS});
        xtb::copy_to_device(a, h_a);
        const std::array<std::size_t, 1> ax0 = {0};

        auto pipeline = [&]() {
            xtb::dist::mean_into<float>(m, a, ax0, total_rows);
            // variance and map both need only the mean: on several GPUs the variance chain runs on the forked stream
            // so that its exchange hides behind the map kernel
            if (world > 1) xtb::check(xtb_fork_begin());
            xtb::dist::variance_into<float>(v, a, m, ax0, total_rows);
            if (world > 1) xtb::check(xtb_fork_end());
            xt::noalias(out) = xt::exp(a - m);
            if (world > 1) xtb::check(xtb_fork_join());
        };
        pipeline();
        pipeline();
        xtb::sync();
        void* graph = nullptr;
        xtb::check(xtb_graph_begin());
        pipeline();
        xtb::check(xtb_graph_end(&graph));
        auto step = [&]() { xtb::check(xtb_graph_launch(graph)); };

        for (int i = 0; i < std::max(warmup, 3); ++i) step();
        xtb::sync();
        max_over_ranks(0.0, world);          // lines the ranks up
        void *e0 = nullptr, *e1 = nullptr;
        xtb::check(xtb_event_create(&e0));
        xtb::check(xtb_event_create(&e1));
        for (int i = 0; i < (world > 1 ? 2 : 0); ++i) step();   // untimed lead-in: the device queues are in step
        xtb::check(xtb_event_record(e0));
        for (int i = 0; i < steps; ++i) step();
        xtb::check(xtb_event_record(e1));
        float ms = 0;
        xtb::check(xtb_event_elapsed_ms(e0, e1, &ms));
        const double dev_ms = max_over_ranks(ms, world) / steps;

        // end to end from pinned host memory
        auto e2e = [&]() {
            xtb::copy_to_device(a, h_a);
            step();
            xtb::copy_to_host(h_m, m);
            xtb::copy_to_host(h_v, v);
            xtb::copy_to_host(h_o, out);
        };
        e2e();
        max_over_ranks(0.0, world);
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < e2e_steps; ++i) e2e();
        const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / e2e_steps;
        const double e2e_ms = max_over_ranks(wall, world);

        // the host result is the device result: every tile of `out` maps the same block
        bool ok = std::equal(h_o.data(), h_o.data() + BLK * COLS, h_o.data() + (rows - BLK) * COLS);
        double checksum = 0;
        for (std::size_t i = 0; i < rows * COLS; i += 4099) checksum += h_o.data()[i];
        double msum = 0, vsum = 0;
        for (std::size_t j = 0; j < COLS; ++j) { msum += h_m.data()[j]; vsum += h_v.data()[j]; }
        const double bytes = 4.0 * double(total_rows) * COLS * 4 + 3.0 * COLS * 4;
        if (rank == 0)
        {
            std::printf("{\"api\": \"include/xtb200/xtensor_b200.hpp (xt::noalias(out) = xt::exp(a - m), xtb::dist::mean_into / variance_into)\", "
                        "\"n_gpus\": %d, \"rows_per_gpu\": %zu, \"steps\": %d, \"ms_per_step\": %.5f, \"value\": %.1f, "
                        "\"e2e_steps\": %d, \"e2e_ms_per_step\": %.3f, \"e2e_value\": %.2f, \"h2d_bytes_per_step\": %zu, "
                        "\"d2h_bytes_per_step\": %zu, \"tiles_identical\": %s, \"result_checksum\": %.9g, \"mean_sum\": %.9g, "
                        "\"variance_mean\": %.9g, \"graph_kernels\": %d, \"peer_memory\": %s}\n",
                        world, rows, steps, dev_ms, bytes / (dev_ms * 1e-3) / 1e9, e2e_steps, e2e_ms, bytes / (e2e_ms * 1e-3) / 1e9,
                        rows * COLS * 4, rows * COLS * 4 + 2 * COLS * 4, ok ? "true" : "false", checksum, msum, vsum / COLS,
                        xtb_graph_kernel_count(graph), cm.peer_memory ? "true" : "false");
        }
        xtb_graph_destroy(graph);
        xtb_event_destroy(e0);
        xtb_event_destroy(e1);
        xtb::sync();
        xtb::dist::finalize();
        return ok ? 0 : 1;
    }
    catch (std::exception& e)
    {
        std::printf("{\"error\": \"%s\"}\n", e.what());
        return 3;
    }
}
