// Drop-in test: the REAL xtensor headers + include/xtb200/xtensor_b200.hpp + libxtb200.so.
// The same expressions are evaluated by xtensor on the CPU (host containers) and by the
// backend (device containers); results must agree bit for bit (+ - * /, integers, integer-
// valued reductions) or within 2 ulp propagated (transcendentals).
// Built here (where /root/reference exists) by tests/cpp/Makefile; the binary travels to the
// GPU box and is run by tests/test_gpu_dropin.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include <xtb200/xtensor_b200.hpp>
#include <xtb200/xtb_npy.hpp>
#include <xtensor/generators/xbuilder.hpp>

static int g_failed = 0;
#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++g_failed; } \
    } while (0)

template <class A, class B> bool same_bits(const A& a, const B& b)
{
    if (a.shape().size() != b.shape().size() || !std::equal(a.shape().begin(), a.shape().end(), b.shape().begin())) return false;
    return std::memcmp(a.data(), b.data(), a.size() * sizeof(typename A::value_type)) == 0;
}

template <class A, class B> double max_abs_diff(const A& a, const B& b)
{
    double m = 0;
    auto ib = b.begin();
    for (auto ia = a.begin(); ia != a.end(); ++ia, ++ib) m = std::max(m, std::fabs(double(*ia) - double(*ib)));
    return m;
}

template <class T, class... S> xt::xarray<T> rnd(unsigned seed, T lo, T hi, S... shape)
{
    std::mt19937_64 gen(seed);
    std::uniform_real_distribution<double> dist((double) lo, (double) hi);
    xt::xarray<T> a = xt::zeros<T>({std::size_t(shape)...});
    for (auto& v : a) v = static_cast<T>(dist(gen));
    return a;
}

int main()
{
    if (xtb_init(0) != 0) { std::printf("xtb_init failed: %s\n", xtb_last_error()); return 2; }

    // cfg1: xt::noalias(c) = a + b on xtensor<double, 1>
    {
        xt::xtensor<double, 1> a = rnd<double>(1, -1, 1, 100003), b = rnd<double>(2, -1, 1, 100003), c;
        xt::noalias(c) = a + b;
        xtb::xtensor<double, 1> da = xtb::to_device(a), db = xtb::to_device(b), dc;
        xt::noalias(dc) = da + db;
        CHECK(same_bits(xtb::to_host(dc), c));
        CHECK(std::strstr(xtb_last_kernel(), "add_f64") != nullptr);
    }
    // cfg2: c = sin(a) * b(1,N,1) + 2.0f * d, value type stays float
    {
        xt::xtensor<float, 3> a = rnd<float>(3, -3, 3, 16, 40, 64), b = rnd<float>(4, 0.5f, 1.5f, 1, 40, 1), d = rnd<float>(5, -3, 3, 16, 40, 64), c;
        xt::noalias(c) = xt::sin(a) * b + 2.0f * d;
        xtb::xtensor<float, 3> da = xtb::to_device(a), db = xtb::to_device(b), dd = xtb::to_device(d), dc;
        auto expr = xt::sin(da) * db + 2.0f * dd;
        static_assert(std::is_same<decltype(expr)::value_type, float>::value, "value type stays float");
        static_assert(xtb::is_b200_expression<decltype(expr)>::value, "one device operand tags the tree");
        xt::noalias(dc) = expr;
        auto hc = xtb::to_host(dc);
        CHECK(hc.shape() == c.shape());
        CHECK(max_abs_diff(hc, c) <= 4e-7 * 4);   // 2 ulp(sin) propagated, |values| < 8
        CHECK(std::strstr(xtb_last_kernel(), "sinmul_axpy_f32") != nullptr);
    }
    // cfg4: out = transpose(a) + view(b, range(0, _, 2), all())
    {
        xt::xtensor<double, 2> a = rnd<double>(7, -1, 1, 96, 96), b = rnd<double>(8, -1, 1, 192, 96), o;
        xt::noalias(o) = xt::transpose(a) + xt::view(b, xt::range(0, xt::placeholders::_, 2), xt::all());
        xtb::xtensor<double, 2> da = xtb::to_device(a), db = xtb::to_device(b), dout;
        xt::noalias(dout) = xt::transpose(da) + xt::view(db, xt::range(0, xt::placeholders::_, 2), xt::all());
        CHECK(same_bits(xtb::to_host(dout), o));
    }
    // cfg5 map, broadcasting a lower-rank operand; benchmark_assign 3.0 * x - 2.0 * y
    {
        xt::xtensor<float, 2> a = rnd<float>(9, -1, 1, 64, 128);
        xt::xtensor<float, 1> m = rnd<float>(10, -0.1f, 0.1f, 128);
        xt::xtensor<float, 2> o;
        xt::noalias(o) = xt::exp(a - m);
        xtb::xtensor<float, 2> da = xtb::to_device(a), dout;
        xtb::xtensor<float, 1> dm = xtb::to_device(m);
        xt::noalias(dout) = xt::exp(da - dm);
        CHECK(max_abs_diff(xtb::to_host(dout), o) <= 1e-6);
        xt::xtensor<double, 2> x = rnd<double>(11, -3, 3, 50, 70), y = rnd<double>(12, -3, 3, 50, 70), r;
        xt::noalias(r) = 3.0 * x - 2.0 * y;
        xtb::xtensor<double, 2> dx = xtb::to_device(x), dy = xtb::to_device(y), dr;
        xt::noalias(dr) = 3.0 * dx - 2.0 * dy;
        CHECK(same_bits(xtb::to_host(dr), r));
    }
    // operator= (alias temporary path), compound assign, xarray, integer promotion, mixed types
    {
        xt::xarray<int> a = xt::arange<int>(0, 60).reshape({3, 4, 5});
        xt::xarray<int> b = xt::arange<int>(1, 21).reshape({4, 5});
        xtb::xarray<int> da = xtb::to_device(a), db = xtb::to_device(b);
        xtb::xarray<int> dc = da * db - da / db + (da % db);
        xt::xarray<int> c = a * b - a / b + (a % b);
        CHECK(same_bits(xtb::to_host(dc), c));
        dc = dc + da;                                   // aliasing assignment: temporary, then move
        c = c + a;
        CHECK(same_bits(xtb::to_host(dc), c));
        xt::noalias(dc) += db;                          // computed assign
        c += b;
        CHECK(same_bits(xtb::to_host(dc), c));
        dc *= 3;                                        // scalar computed assign (core/xassign.hpp:525-537)
        c *= 3;
        dc -= 7;
        c -= 7;
        CHECK(same_bits(xtb::to_host(dc), c));
        xt::xarray<double> q = rnd<double>(21, -4, 4, 3, 4, 5);
        xtb::xarray<double> dq = xtb::to_device(q);
        xt::noalias(dq) += 3.123;                       // benchmark_assign.cpp assign_x_scalar_computed
        q += 3.123;
        dq /= 1.5;
        q /= 1.5;
        CHECK(same_bits(xtb::to_host(dq), q));
        xt::xarray<float> f = rnd<float>(13, -4, 4, 3, 4, 5);
        xtb::xarray<float> df = xtb::to_device(f);
        auto mixed = 2.0 * df + da;                     // double * float + int -> double
        static_assert(std::is_same<decltype(mixed)::value_type, double>::value, "C++ promotion");
        xtb::xarray<double> dm = mixed;
        xt::xarray<double> hm = 2.0 * f + a;
        CHECK(same_bits(xtb::to_host(dm), hm));
        xtb::xarray<double> dw = xt::where(df > 0.5f, xt::sqrt(xt::abs(df)), xt::square(df));
        xt::xarray<double> hw = xt::where(f > 0.5f, xt::sqrt(xt::abs(f)), xt::square(f));
        CHECK(max_abs_diff(xtb::to_host(dw), hw) == 0.0);
    }
    // reducers: lazy xreducer assigned to a device container; mean; nested in an expression
    {
        xt::xarray<float> a = xt::round(rnd<float>(14, -8, 8, 24, 30, 16));    // integer valued: exact in any order
        xtb::xarray<float> da = xtb::to_device(a);
        xtb::xarray<float> s0 = xt::sum(da, {0}), s2 = xt::sum(da, {2}), s02 = xt::sum(da, {0, 2}), mx = xt::amax(da, {1});
        xt::xarray<float> h0 = xt::sum(a, {0}), h2 = xt::sum(a, {2}), h02 = xt::sum(a, {0, 2}), hmx = xt::amax(a, {1});
        CHECK(same_bits(xtb::to_host(s0), h0));
        CHECK(same_bits(xtb::to_host(s2), h2));
        CHECK(same_bits(xtb::to_host(s02), h02));
        CHECK(same_bits(xtb::to_host(mx), hmx));
        auto me = xt::mean(da, {0});
        static_assert(std::is_same<decltype(me)::value_type, double>::value, "mean of float is double");
        xtb::xarray<double> dmean = me;
        xt::xarray<double> hmean = xt::mean(a, {0});
        CHECK(same_bits(xtb::to_host(dmean), hmean));
        xtb::xarray<float> ks = xt::sum(da, {1}, xt::keep_dims);
        xt::xarray<float> hks = xt::sum(a, {1}, xt::keep_dims);
        CHECK(same_bits(xtb::to_host(ks), hks));
        xtb::xarray<float> centered = da - xt::sum(da, {0}) / 24.0f;            // reducer nested in a tree
        xt::xarray<float> hcent = a - xt::sum(a, {0}) / 24.0f;
        CHECK(same_bits(xtb::to_host(centered), hcent));
    }
    // immediate evaluation strategy and accumulators (overloads selected by the expression tag)
    {
        xt::xarray<double> a = xt::round(rnd<double>(15, -8, 8, 12, 20, 9));
        xtb::xarray<double> da = xtb::to_device(a);
        auto si = xt::sum(da, {0, 1}, xt::evaluation_strategy::immediate);
        xt::xarray<double> hsi = xt::sum(a, {0, 1}, xt::evaluation_strategy::immediate);
        CHECK(same_bits(xtb::to_host(si), hsi));
        auto mi = xt::amin(da, {2}, xt::keep_dims | xt::evaluation_strategy::immediate);
        xt::xarray<double> hmi = xt::amin(a, {2}, xt::keep_dims | xt::evaluation_strategy::immediate);
        CHECK(same_bits(xtb::to_host(mi), hmi));
        xt::xarray<double> hvar = xt::variance(a, {0});                 // two-pass, uses eval(mean(immediate))
        xtb::xarray<double> dvar = xt::variance(da, {0});
        CHECK(max_abs_diff(xtb::to_host(dvar), hvar) <= 1e-12 * 64);
        auto c1 = xt::cumsum(da, 1);
        xt::xarray<double> hc1 = xt::cumsum(a, 1);
        CHECK(same_bits(xtb::to_host(c1), hc1));
        auto cf = xt::cumsum(da);
        xt::xarray<double> hcf = xt::cumsum(a);
        CHECK(same_bits(xtb::to_host(cf), hcf));
        xt::xarray<short> sh = {short(1), short(2), short(3), short(4)};  // test_xaccumulator.cpp:22-34
        auto cs = xt::cumsum(xtb::xarray<short>(xtb::to_device(sh)));
        static_assert(std::is_same<decltype(cs)::value_type, int>::value, "short -> int promotion");
        xt::xarray<int> expect = {1, 3, 6, 10};
        CHECK(same_bits(xtb::to_host(cs), expect));
        auto cp = xt::cumprod(da * 0.5 + 1.0, 2);                           // accumulator over an expression
        xt::xarray<double> hcp = xt::cumprod(a * 0.5 + 1.0, 2);
        CHECK(max_abs_diff(xtb::to_host(cp), hcp) <= 1e-9 * std::fabs(hcp(0, 0, 8)) + 1e-6);
    }
    // nan-skipping reducers (core/xmath.hpp:2365-2473, test/test_xnan_functions.cpp): nansum / nanprod lower to
    // the sum / prod kernels with the NaN replacement fused into the map program; nanmean composes them
    {
        xt::xarray<double> a = xt::round(rnd<double>(16, -3, 3, 10, 12, 7));
        a(0, 0, 0) = std::nan("");
        a(3, 5, 2) = std::nan("");
        for (std::size_t i = 0; i < 10; ++i) a(i, 4, 6) = std::nan("");     // an all-NaN lane along axis 0
        xtb::xarray<double> da = xtb::to_device(a);
        xtb::xarray<double> ns0 = xt::nansum(da, {0}), ns12 = xt::nansum(da, {1, 2}), np2 = xt::nanprod(da, {2});
        xt::xarray<double> hs0 = xt::nansum(a, {0}), hs12 = xt::nansum(a, {1, 2}), hp2 = xt::nanprod(a, {2});
        CHECK(same_bits(xtb::to_host(ns0), hs0));
        CHECK(same_bits(xtb::to_host(ns12), hs12));
        CHECK(same_bits(xtb::to_host(np2), hp2));
        auto nsi = xt::nansum(da, {0, 2}, xt::evaluation_strategy::immediate);
        xt::xarray<double> hnsi = xt::nansum(a, {0, 2}, xt::evaluation_strategy::immediate);
        CHECK(same_bits(xtb::to_host(nsi), hnsi));
    }
    // adaptors over foreign device memory (xt::adapt with no_ownership, containers/xadapt.hpp:105-215): operands
    // and destination are raw device pointers, nothing is copied or owned
    {
        xt::xarray<float> a = rnd<float>(41, -2, 2, 6, 40), b = rnd<float>(42, -2, 2, 40);
        float *pa = nullptr, *pb = nullptr, *po = nullptr;
        xtb::check(xtb_malloc(a.size() * sizeof(float), reinterpret_cast<void**>(&pa)));
        xtb::check(xtb_malloc(b.size() * sizeof(float), reinterpret_cast<void**>(&pb)));
        xtb::check(xtb_malloc(a.size() * sizeof(float), reinterpret_cast<void**>(&po)));
        xtb::check(xtb_memcpy(pa, a.data(), a.size() * sizeof(float), XTB_H2D));
        xtb::check(xtb_memcpy(pb, b.data(), b.size() * sizeof(float), XTB_H2D));
        auto da = xtb::adapt(pa, {6, 40});
        auto db = xtb::adapt(pb, {40});
        auto dout = xtb::adapt(po, {6, 40});
        xt::noalias(dout) = da * db + 2.0f;                               // broadcast, written straight into po
        xt::xarray<float> expect = a * b + 2.0f, got = xt::zeros<float>({6, 40});
        xtb::check(xtb_memcpy(got.data(), po, got.size() * sizeof(float), XTB_D2H));
        xtb::check(xtb_sync());
        CHECK(same_bits(got, expect));
        std::vector<std::size_t> tshape = {40, 6};
        std::vector<std::ptrdiff_t> tstrides = {1, 40};                   // the same memory seen transposed
        auto dat = xtb::adapt(pa, tshape, tstrides);
        xtb::xarray<float> tsum = xt::sum(dat, {1});
        xt::xarray<float> hsum = xt::sum(xt::transpose(a), {1});
        CHECK(max_abs_diff(xtb::to_host(tsum), hsum) <= 1e-5);
        xtb::xarray<float> owned = da + 1.0f;                             // adaptor operand, owning destination
        xt::xarray<float> howned = a + 1.0f;
        CHECK(same_bits(xtb::to_host(owned), howned));
        dout = xt::sqrt(xt::abs(da));                                     // aliasing path: temporary, then copied into po
        xt::xarray<float> expect2 = xt::sqrt(xt::abs(a));
        xtb::check(xtb_memcpy(got.data(), po, got.size() * sizeof(float), XTB_D2H));
        xtb::check(xtb_sync());
        CHECK(same_bits(got, expect2));
        xtb::check(xtb_free(pa));
        xtb::check(xtb_free(pb));
        xtb::check(xtb_free(po));
    }
    // view on the left-hand side (xview_semantic, core/xsemantic.hpp:726-796): strided store and
    // broadcasting into a view (test_strided_assign.cpp:178-196, test_extended_broadcast_view.cpp:811-1033)
    {
        xt::xarray<int> buf = xt::ones<int>({4, 6}) * -1;
        xt::xarray<int> src = xt::arange<int>(1, 9).reshape({2, 4});
        xtb::xarray<int> dbuf = xtb::to_device(buf), dsrc = xtb::to_device(src);
        xt::noalias(xt::view(dbuf, xt::range(1, 3), xt::range(2, 6))) = dsrc;
        xt::noalias(xt::view(buf, xt::range(1, 3), xt::range(2, 6))) = src;
        CHECK(same_bits(xtb::to_host(dbuf), buf));
        xt::xarray<int> row = xt::arange<int>(10, 16);
        xtb::xarray<int> drow = xtb::to_device(row);
        xt::noalias(xt::view(dbuf, xt::range(0, 4, 2), xt::all())) = drow * 2;   // broadcast a row into every other row
        xt::view(buf, xt::range(0, 4, 2), xt::all()) = row * 2;
        CHECK(same_bits(xtb::to_host(dbuf), buf));
        xt::noalias(xt::view(dbuf, 3, xt::all())) += drow;                   // computed assign on a view
        xt::noalias(xt::view(buf, 3, xt::all())) += row;
        CHECK(same_bits(xtb::to_host(dbuf), buf));
        xt::xarray<double> a = rnd<double>(31, -1, 1, 6, 5), out = xt::zeros<double>({5, 6});
        xtb::xarray<double> da = xtb::to_device(a), dout = xtb::to_device(out);
        xt::noalias(xt::transpose(dout)) = da * 2.0;                         // strided (transposed) destination
        xt::noalias(xt::transpose(out)) = a * 2.0;
        CHECK(same_bits(xtb::to_host(dout), out));
    }
    // broadcast error is raised by xtensor's own shape logic before any kernel is launched
    {
        xtb::xtensor<float, 2> da = xtb::to_device(xt::xtensor<float, 2>(xt::ones<float>({3, 4})));
        xtb::xtensor<float, 1> db = xtb::to_device(xt::xtensor<float, 1>(xt::ones<float>({5})));
        bool threw = false;
        try { xtb::xtensor<float, 2> dc = da + db; } catch (const xt::broadcast_error&) { threw = true; }
        CHECK(threw);
    }
    // .npy fixtures <-> device containers (io/xnpy.hpp:740-800 through xtb::load_npy / xtb::dump_npy)
    {
        xt::xarray<float> a = rnd<float>(41, -2, 2, 37, 19);
        const std::string in = "/tmp/xtb200_dropin_in.npy", out = "/tmp/xtb200_dropin_out.npy";
        xt::dump_npy(in, a);
        xtb::xarray<float> da = xtb::load_npy<float>(in);
        CHECK(same_bits(xtb::to_host(da), a));
        xtb::dump_npy(out, da * 2.0f + 1.0f);                // a lazy device expression is evaluated on the device first
        xt::xarray<float> back = xt::load_npy<float>(out);
        xt::xarray<float> want = a * 2.0f + 1.0f;
        CHECK(same_bits(back, want));
        std::remove(in.c_str());
        std::remove(out.c_str());
    }
    xtb::sync();
    std::printf("%s (%d failures, %lld kernel launches)\n", g_failed ? "FAILED" : "OK", g_failed, (long long) xtb_launch_count(0));
    return g_failed ? 1 : 0;
}
