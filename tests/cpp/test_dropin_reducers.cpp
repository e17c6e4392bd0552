// Drop-in test, second part: the reducers whose functors are lambdas or index-carrying -- nan-family, counts,
// minmax, argmin / argmax, the norms, average -- plus xt::initial on narrow integers, generator leaves, the
// fused mean finalize and the pinned-host end-to-end call.  Same method as test_dropin.cpp: every expression
// is evaluated by the REAL xtensor on host containers and by the backend on device containers.
// References: core/xmath.hpp:2195-2228 (minmax), :2307-2860 (nan functions, counts), :1925-2010 (average),
// misc/xsort.hpp:1150-1300 (argmin / argmax), reducers/xnorm.hpp:369-620, test/test_xsort.cpp:217-281,
// test/test_xnan_functions.cpp, test/test_xnorm.cpp.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>

#include <xtb200/xtensor_b200.hpp>
#include <xtensor/generators/xbuilder.hpp>

static int g_failed = 0;
#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++g_failed; } \
    } while (0)

template <class A, class B> bool eq_shape(const A& a, const B& b)
{
    return a.shape().size() == b.shape().size() && std::equal(a.shape().begin(), a.shape().end(), b.shape().begin());
}
// bit-identical, NaN == NaN
template <class A, class B> bool same_bits(const A& a, const B& b)
{
    if (!eq_shape(a, b)) return false;
    using T = typename A::value_type;
    auto ib = b.begin();
    for (auto ia = a.begin(); ia != a.end(); ++ia, ++ib)
    {
        const T x = *ia, y = static_cast<T>(*ib);
        if constexpr (std::is_floating_point<T>::value)
        {
            if (std::isnan(x) && std::isnan(y)) continue;
        }
        if (std::memcmp(&x, &y, sizeof(T)) != 0) return false;
    }
    return true;
}
template <class A, class B> double max_rel_diff(const A& a, const B& b)
{
    if (!eq_shape(a, b)) return 1e300;
    double m = 0;
    auto ib = b.begin();
    for (auto ia = a.begin(); ia != a.end(); ++ia, ++ib)
    {
        const double x = double(*ia), y = double(*ib);
        if (std::isnan(x) && std::isnan(y)) continue;
        if (std::isnan(x) != std::isnan(y)) return 1e300;
        m = std::max(m, std::fabs(x - y) / std::max(1e-300, std::max(std::fabs(x), std::fabs(y))));
    }
    return m;
}

template <class T> xt::xarray<T> rnd_int(unsigned seed, int lo, int hi, std::vector<std::size_t> shape)
{
    std::mt19937_64 gen(seed);
    std::uniform_int_distribution<int> dist(lo, hi);
    xt::xarray<T> a = xt::zeros<T>(shape);
    for (auto& v : a) v = static_cast<T>(dist(gen));
    return a;
}
template <class T> void sprinkle_nan(xt::xarray<T>& a, unsigned seed, double frac)
{
    std::mt19937_64 gen(seed);
    std::uniform_real_distribution<double> u(0, 1);
    for (auto& v : a) if (u(gen) < frac) v = std::numeric_limits<T>::quiet_NaN();
}

template <class T> void nan_family()
{
    xt::xarray<T> a = rnd_int<T>(31, -3, 3, {6, 5, 7});      // small integers: sums / products exact in any order
    sprinkle_nan(a, 32, 0.2);
    xt::view(a, 2, xt::all(), 3) = std::numeric_limits<T>::quiet_NaN();    // all-NaN lanes along every axis
    xt::view(a, xt::all(), 4, 6) = std::numeric_limits<T>::quiet_NaN();
    xt::view(a, 5, 1, xt::all()) = std::numeric_limits<T>::quiet_NaN();
    xtb::xarray<T> da = xtb::to_device(a);
    const std::vector<std::vector<std::size_t>> axes = {{0}, {1}, {2}, {0, 1}, {1, 2}, {0, 2}, {0, 1, 2}};
    for (const auto& ax : axes)
    {
        { xtb::xarray<T> d = xt::nansum(da, ax); xt::xarray<T> h = xt::nansum(a, ax); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<T> d = xt::nanprod(da, ax); xt::xarray<T> h = xt::nanprod(a, ax); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<T> d = xt::nanmin(da, ax); xt::xarray<T> h = xt::nanmin(a, ax); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<T> d = xt::nanmax(da, ax); xt::xarray<T> h = xt::nanmax(a, ax); CHECK(same_bits(xtb::to_host(d), h)); }
        {
            auto dc = xt::count_nonzero(da, ax);
            static_assert(std::is_same<typename decltype(dc)::value_type, std::size_t>::value, "count type");
            xtb::xarray<std::size_t> d = dc;
            xt::xarray<std::size_t> h = xt::count_nonzero(a, ax);
            CHECK(same_bits(xtb::to_host(d), h));
        }
        { xtb::xarray<std::size_t> d = xt::count_nonnan(da, ax); xt::xarray<std::size_t> h = xt::count_nonnan(a, ax); CHECK(same_bits(xtb::to_host(d), h)); }
        {
            auto dm = xt::nanmean(da, ax);
            static_assert(std::is_same<typename decltype(dm)::value_type, double>::value, "nanmean defaults to double");
            xtb::xarray<double> d = dm;
            xt::xarray<double> h = xt::nanmean(a, ax);
            CHECK(same_bits(xtb::to_host(d), h));
        }
        { xtb::xarray<T> d = xt::nanmean<T>(da, ax); xt::xarray<T> h = xt::nanmean<T>(a, ax); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<double> d = xt::nanvar(da, ax); xt::xarray<double> h = xt::nanvar(a, ax); CHECK(max_rel_diff(xtb::to_host(d), h) <= 1e-12); }
        { xtb::xarray<double> d = xt::nanstd(da, ax); xt::xarray<double> h = xt::nanstd(a, ax); CHECK(max_rel_diff(xtb::to_host(d), h) <= 1e-12); }
    }
    // immediate strategy goes through reduce_immediate
    {
        auto d = xt::nanmin(da, {1}, xt::evaluation_strategy::immediate);
        xt::xarray<T> h = xt::nanmin(a, {1}, xt::evaluation_strategy::immediate);
        CHECK(same_bits(xtb::to_host(d), h));
        auto c = xt::count_nonzero(da, {0, 2}, xt::evaluation_strategy::immediate);
        xt::xarray<std::size_t> hc = xt::count_nonzero(a, {0, 2}, xt::evaluation_strategy::immediate);
        CHECK(same_bits(xtb::to_host(c), hc));
    }
}

template <class T> void arg_family(bool with_nan)
{
    xt::xarray<T> a = rnd_int<T>(41, -5, 5, {7, 6, 9});       // many ties: the FIRST extreme must win
    if constexpr (std::is_floating_point<T>::value)
    {
        if (with_nan)
        {
            sprinkle_nan(a, 42, 0.15);
            a(0, 0, 0) = std::numeric_limits<T>::quiet_NaN();                  // NaN first: sticks (index 0)
            xt::view(a, 3, xt::all(), 4) = std::numeric_limits<T>::quiet_NaN();    // all-NaN lane
        }
    }
    xtb::xarray<T> da = xtb::to_device(a);
    for (std::ptrdiff_t ax : {0, 1, 2, -1})
    {
        { auto d = xt::argmin(da, ax); xt::xarray<std::size_t> h = xt::argmin(a, ax); CHECK(same_bits(xtb::to_host(d), h)); }
        { auto d = xt::argmax(da, ax); xt::xarray<std::size_t> h = xt::argmax(a, ax); CHECK(same_bits(xtb::to_host(d), h)); }
    }
    { auto d = xt::argmin(da); xt::xarray<std::size_t> h = xt::argmin(a); CHECK(same_bits(xtb::to_host(d), h)); }
    { auto d = xt::argmax(da); xt::xarray<std::size_t> h = xt::argmax(a); CHECK(same_bits(xtb::to_host(d), h)); }
    // an expression operand (evaluated first, as the reference does) and a strided view, flat
    { auto d = xt::argmax(da * da - da, 1); xt::xarray<std::size_t> h = xt::argmax(a * a - a, 1); CHECK(same_bits(xtb::to_host(d), h)); }
    {
        auto dv = xt::view(da, xt::range(1, 6, 2), xt::all(), xt::range(0, 8, 3));
        auto hv = xt::view(a, xt::range(1, 6, 2), xt::all(), xt::range(0, 8, 3));
        auto d = xt::argmin(dv);
        xt::xarray<std::size_t> h = xt::argmin(hv);
        CHECK(same_bits(xtb::to_host(d), h));
        auto d2 = xt::argmax(dv, 2);
        xt::xarray<std::size_t> h2 = xt::argmax(hv, 2);
        CHECK(same_bits(xtb::to_host(d2), h2));
    }
}

// the norm overloads taking an axes CONTAINER are ambiguous in the reference itself (norm(E&&, X&&, EVS) vs
// norm(E&&, EVS)); its tests pass brace lists (test/test_xnorm.cpp), which select the C-array overload
#define NORM_CASE(FN, CMP, ...)                                                             \
    {                                                                                       \
        auto dn = FN(da, __VA_ARGS__);                                                      \
        using R = typename decltype(dn)::value_type;                                        \
        xtb::xarray<R> d = dn;                                                              \
        xt::xarray<R> h = FN(a, __VA_ARGS__);                                               \
        CHECK(CMP);                                                                         \
    }
#define NORM_AXES(...)                                                                                          \
    NORM_CASE(xt::norm_l0, same_bits(xtb::to_host(d), h), __VA_ARGS__)                                          \
    NORM_CASE(xt::norm_l1, same_bits(xtb::to_host(d), h), __VA_ARGS__)                                          \
    NORM_CASE(xt::norm_sq, same_bits(xtb::to_host(d), h), __VA_ARGS__)                                          \
    NORM_CASE(xt::norm_linf, same_bits(xtb::to_host(d), h), __VA_ARGS__)                                        \
    NORM_CASE(xt::norm_l2, max_rel_diff(xtb::to_host(d), h) <= 1e-15, __VA_ARGS__)

template <class T> void norm_family()
{
    xt::xarray<T> a = rnd_int<T>(51, std::is_signed<T>::value ? -6 : 0, 6, {5, 8, 6});
    xtb::xarray<T> da = xtb::to_device(a);
    NORM_AXES({0})
    NORM_AXES({2})
    NORM_AXES({0, 2})
    NORM_AXES({0, 1, 2})
    // float operands raise to the power in float (real_promote_type_t<float>): 2 ulp(fp32) per term vs glibc's powf
    const double ptol = std::is_same<T, float>::value ? 3e-7 : 1e-13;
    NORM_CASE(xt::norm_lp_to_p, max_rel_diff(xtb::to_host(d), h) <= ptol, 3.0, {0})
    NORM_CASE(xt::norm_lp_to_p, max_rel_diff(xtb::to_host(d), h) <= ptol, 2.5, {1, 2})
    NORM_CASE(xt::norm_lp, max_rel_diff(xtb::to_host(d), h) <= ptol, 1.5, {2})
    NORM_CASE(xt::norm_lp, max_rel_diff(xtb::to_host(d), h) <= ptol, 4.0, {0, 1, 2})
    {
        auto dn = xt::norm_l1(da);
        using R = typename decltype(dn)::value_type;
        xtb::xarray<R> d = dn;
        xt::xarray<R> h = xt::norm_l1(a);
        CHECK(same_bits(xtb::to_host(d), h));
    }
}

int main()
{
    if (xtb_init(0) != 0) { std::printf("xtb_init failed: %s\n", xtb_last_error()); return 2; }

    nan_family<float>();
    nan_family<double>();

    // integer operands of the nan / count reducers (isnan(int) is false; counts of a bool expression)
    {
        xt::xarray<int> i = rnd_int<int>(33, -2, 2, {6, 5, 7});
        xtb::xarray<int> di = xtb::to_device(i);
        xtb::xarray<std::size_t> d = xt::count_nonzero(di, {1}); xt::xarray<std::size_t> h = xt::count_nonzero(i, {1});
        CHECK(same_bits(xtb::to_host(d), h));
        xtb::xarray<std::size_t> d2 = xt::count_nonzero(di > 0); xt::xarray<std::size_t> h2 = xt::count_nonzero(i > 0);
        CHECK(same_bits(xtb::to_host(d2), h2));
        xtb::xarray<int> dm = xt::nanmin(di, {0, 2}); xt::xarray<int> hm = xt::nanmin(i, {0, 2});
        CHECK(same_bits(xtb::to_host(dm), hm));
    }

    // minmax (core/xmath.hpp:2195-2228): lazy and immediate, float with NaN (std::min / std::max skip them) and int
    {
        xt::xarray<float> a = rnd_int<float>(34, -50, 50, {9, 11, 5});
        xtb::xarray<float> da = xtb::to_device(a);
        xtb::xtensor<std::array<float, 2>, 0> d = xt::minmax(da);
        xt::xtensor<std::array<float, 2>, 0> h = xt::minmax(a);
        auto hd = xtb::to_host(d);
        CHECK(hd()[0] == h()[0] && hd()[1] == h()[1]);
        sprinkle_nan(a, 35, 0.3);
        da = xtb::to_device(a);
        auto di = xt::minmax(da, xt::evaluation_strategy::immediate);
        auto hi = xt::minmax(a, xt::evaluation_strategy::immediate);
        auto hdi = xtb::to_host(di);
        CHECK(hdi()[0] == hi()[0] && hdi()[1] == hi()[1]);
        xt::xarray<short> s = rnd_int<short>(36, -3000, 3000, {40, 33});
        xtb::xarray<short> ds = xtb::to_device(s);
        xtb::xtensor<std::array<short, 2>, 0> dms = xt::minmax(ds);
        xt::xtensor<std::array<short, 2>, 0> hms = xt::minmax(s);
        auto hdms = xtb::to_host(dms);
        CHECK(hdms()[0] == hms()[0] && hdms()[1] == hms()[1]);
    }

    // argmin / argmax: the reference's own cases (test/test_xsort.cpp:217-281) ...
    {
        xt::xarray<double> a = {{5, 3, 1}, {4, 4, 4}};
        xtb::xarray<double> da = xtb::to_device(a);
        xt::xarray<std::size_t> ex = std::size_t(2);
        xt::xtensor<std::size_t, 1> ex_2 = {1, 0, 0}, ex_3 = {2, 0}, ex_mx0 = {0, 1, 1}, ex_mx1 = {0, 0};
        CHECK(same_bits(xtb::to_host(xt::argmin(da)), ex));
        CHECK(same_bits(xtb::to_host(xt::argmin(da, 0)), ex_2));
        CHECK(same_bits(xtb::to_host(xt::argmin(da, 1)), ex_3));
        CHECK(xtb::to_host(xt::argmax(da))() == 0ul);
        CHECK(same_bits(xtb::to_host(xt::argmax(da, 0)), ex_mx0));
        CHECK(same_bits(xtb::to_host(xt::argmax(da, 1)), ex_mx1));
        xt::xarray<double> b = {1, 3, 4, -100};
        xtb::xarray<double> db = xtb::to_device(b);
        CHECK(xtb::to_host(xt::argmin(db))() == 3ul);
        CHECK(xtb::to_host(xt::argmin(db, 0))() == 3ul);
        xt::xtensor<int, 3> c = {{{1, 2, 3, 4}}, {{4, 3, 2, 1}}};
        xtb::xtensor<int, 3> dc = xtb::to_device(c);
        xt::xtensor<std::size_t, 2> ex_4 = {{3}, {0}}, ex_5 = {{1, 1, 0, 0}}, ex_6 = {{0, 0, 0, 0}, {0, 0, 0, 0}};
        CHECK(same_bits(xtb::to_host(xt::argmax(dc, 2)), ex_4));
        CHECK(same_bits(xtb::to_host(xt::argmax(dc, 0)), ex_5));
        CHECK(same_bits(xtb::to_host(xt::argmax(dc, 1)), ex_6));
        xt::xarray<double> ya = {1, 0, 3, 2, 2}, d3 = {0, 1, 0};
        CHECK(xtb::to_host(xt::argmin(xtb::xarray<double>(xtb::to_device(ya))))() == 1ul);
        CHECK(xtb::to_host(xt::argmax(xtb::xarray<double>(xtb::to_device(ya)), 0))() == 2ul);
        CHECK(xtb::to_host(xt::argmax(xtb::xarray<double>(xtb::to_device(d3))))() == 1ul);   // xtensor#2568
    }
    // ... and sweeps against the reference over dtypes, axes, ties and NaNs
    arg_family<float>(false);
    arg_family<float>(true);
    arg_family<double>(true);
    arg_family<int>(false);
    arg_family<short>(false);
    arg_family<unsigned char>(false);
    arg_family<long long>(false);

    // norms (reducers/xnorm.hpp)
    norm_family<float>();
    norm_family<double>();
    norm_family<int>();
    norm_family<unsigned short>();

    // average (core/xmath.hpp:1925-2010): weights along one axis, of the full shape, whole array
    {
        xt::xarray<double> a = rnd_int<double>(61, -5, 5, {4, 6, 5}), w1 = rnd_int<double>(62, 1, 4, {6}), wf = rnd_int<double>(63, 1, 4, {4, 6, 5});
        xtb::xarray<double> da = xtb::to_device(a), dw1 = xtb::to_device(w1), dwf = xtb::to_device(wf);
        { xtb::xarray<double> d = xt::average(da, dw1, {1}); xt::xarray<double> h = xt::average(a, w1, {1}); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<double> d = xt::average(da, dwf, {0, 2}); xt::xarray<double> h = xt::average(a, wf, {0, 2}); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<double> d = xt::average(da, dwf); xt::xarray<double> h = xt::average(a, wf); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<double> d = xt::average(da); xt::xarray<double> h = xt::average(a); CHECK(same_bits(xtb::to_host(d), h)); }
    }

    // xt::initial on narrow integers (the accumulator's register type is int): amax / amin / sum
    {
        xt::xarray<signed char> a = rnd_int<signed char>(71, -100, 100, {12, 9});
        xt::xarray<short> s = rnd_int<short>(72, -30000, 30000, {12, 9});
        xtb::xarray<signed char> da = xtb::to_device(a);
        xtb::xarray<short> ds = xtb::to_device(s);
        { xtb::xarray<signed char> d = xt::amax(da, {0}, xt::initial((signed char) 17)); xt::xarray<signed char> h = xt::amax(a, {0}, xt::initial((signed char) 17)); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<signed char> d = xt::amin(da, {1}, xt::initial((signed char) -5)); xt::xarray<signed char> h = xt::amin(a, {1}, xt::initial((signed char) -5)); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<short> d = xt::amax(ds, {0}, xt::initial((short) 12345)); xt::xarray<short> h = xt::amax(s, {0}, xt::initial((short) 12345)); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<short> d = xt::amin(ds, {0, 1}, xt::initial((short) -31000)); xt::xarray<short> h = xt::amin(s, {0, 1}, xt::initial((short) -31000)); CHECK(same_bits(xtb::to_host(d), h)); }
        {
            auto dsum = xt::sum(ds, {1}, xt::initial(100));
            using R = typename decltype(dsum)::value_type;
            xtb::xarray<R> d = dsum; xt::xarray<R> h = xt::sum(s, {1}, xt::initial(100)); CHECK(same_bits(xtb::to_host(d), h));
            auto di = xt::amax(ds, {1}, xt::initial((short) 5) | xt::evaluation_strategy::immediate);
            xt::xarray<short> hi = xt::amax(s, {1}, xt::initial((short) 5) | xt::evaluation_strategy::immediate);
            CHECK(same_bits(xtb::to_host(di), hi));
        }
    }

    // generator leaves: ones / zeros are scalar broadcasts, arange / linspace are host fills uploaded once
    {
        xt::xarray<float> a = rnd_int<float>(81, -9, 9, {6, 10});
        xtb::xarray<float> da = xtb::to_device(a);
        { xtb::xarray<float> d = da + xt::ones<float>({6, 10}); xt::xarray<float> h = a + xt::ones<float>({6, 10}); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<float> d = da * xt::arange<float>(0.f, 10.f); xt::xarray<float> h = a * xt::arange<float>(0.f, 10.f); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<double> d = da - xt::linspace<double>(0., 1., 10); xt::xarray<double> h = a - xt::linspace<double>(0., 1., 10); CHECK(same_bits(xtb::to_host(d), h)); }
    }

    // mean / variance / stddev: the division is fused into the reduction's last store (one kernel, no temporary)
    {
        xt::xarray<float> a = rnd_int<float>(91, -8, 8, {96, 130});
        xtb::xarray<float> da = xtb::to_device(a);
        xtb_launch_count(1);
        xtb::xarray<double> dm = xt::mean(da, {0});
        CHECK(xtb_launch_count(0) <= 2 && std::strstr(xtb_last_kernel(), "k_reduce") != nullptr);   // reduce (+ merge of split rows): no separate divide kernel
        xt::xarray<double> hm = xt::mean(a, {0});
        CHECK(same_bits(xtb::to_host(dm), hm));
        xtb::xarray<float> dmf = xt::mean<float>(da, {1}); xt::xarray<float> hmf = xt::mean<float>(a, {1});
        CHECK(same_bits(xtb::to_host(dmf), hmf));
        xtb::xarray<float> dv = xt::variance<float>(da, {0}); xt::xarray<float> hv = xt::variance<float>(a, {0});
        CHECK(max_rel_diff(xtb::to_host(dv), hv) <= 1e-6);
        xtb::xarray<double> dsd = xt::stddev(da, {1}); xt::xarray<double> hsd = xt::stddev(a, {1});
        CHECK(max_rel_diff(xtb::to_host(dsd), hsd) <= 1e-12);
    }

    // end to end from pinned HOST containers: one pipelined call (H2D | kernel | D2H), no CPU evaluation
    {
        xtb::pinned_xtensor<float, 3> a, d, c;
        xtb::pinned_xtensor<float, 3> b;
        a.resize({64, 40, 64}); d.resize({64, 40, 64}); b.resize({1, 40, 1});
        std::mt19937_64 gen(7);
        std::uniform_real_distribution<float> u(-3.f, 3.f);
        for (auto& v : a) v = u(gen);
        for (auto& v : d) v = u(gen);
        for (auto& v : b) v = 1.f + 0.1f * u(gen);
        xtb::assign_host(c, xt::sin(a) * b + 2.0f * d, 1 << 20);      // small chunks: several pipeline stages
        xt::xtensor<float, 3> ha = a, hb = b, hd = d, hc;
        xt::noalias(hc) = xt::sin(ha) * hb + 2.0f * hd;
        CHECK(eq_shape(c, hc));
        double m = 0;
        auto ih = hc.begin();
        for (auto ic = c.begin(); ic != c.end(); ++ic, ++ih) m = std::max(m, std::fabs(double(*ic) - double(*ih)));
        CHECK(m <= 4e-7 * 4);
        // and it agrees bit for bit with the device-resident path
        xtb::xtensor<float, 3> da = xtb::to_device(ha), db = xtb::to_device(hb), dd = xtb::to_device(hd), dc;
        xt::noalias(dc) = xt::sin(da) * db + 2.0f * dd;
        CHECK(same_bits(xtb::to_host(dc), c));
    }

    // shapes the planner rewrites into two passes (narrow: < 1024 outputs; mixed: outer + innermost axis) and odd
    // extents (scalar-access kernels), through the unchanged xtensor calls; integer-valued data: exact in any order
    {
        xt::xarray<float> a = rnd_int<float>(51, -4, 4, {70001, 20});                 // narrow, remainder slice (70001 is prime-ish)
        xtb::xarray<float> da = xtb::to_device(a);
        { xtb::xarray<float> d = xt::sum(da, {0}); xt::xarray<float> h = xt::sum(a, {0}); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<float> d = xt::amax(da, {0}); xt::xarray<float> h = xt::amax(a, {0}); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<double> d = xt::mean(da, {0}); xt::xarray<double> h = xt::mean(a, {0}); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<float> d = xt::sum(da, {0}, xt::keep_dims | xt::initial(7.0f)); xt::xarray<float> h = xt::sum(a, {0}, xt::keep_dims | xt::initial(7.0f));
          CHECK(same_bits(xtb::to_host(d), h)); }
        xt::xarray<double> b = rnd_int<double>(52, -4, 4, {300, 17, 257});              // mixed: axes {0, 2}, kept dim between
        xtb::xarray<double> db = xtb::to_device(b);
        { xtb::xarray<double> d = xt::sum(db, {0, 2}); xt::xarray<double> h = xt::sum(b, {0, 2}); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<std::size_t> d = xt::count_nonzero(db, {0, 2}); xt::xarray<std::size_t> h = xt::count_nonzero(b, {0, 2}); CHECK(same_bits(xtb::to_host(d), h)); }
        xt::xarray<int> c = rnd_int<int>(53, -9, 9, {2047, 1022});                       // odd row pitch: 4088 bytes
        xtb::xarray<int> dc = xtb::to_device(c);
        { xtb::xarray<int> d = xt::sum(dc, {0}); xt::xarray<int> h = xt::sum(c, {0}); CHECK(same_bits(xtb::to_host(d), h)); }
        { xtb::xarray<int> d = xt::amin(dc, {1}); xt::xarray<int> h = xt::amin(c, {1}); CHECK(same_bits(xtb::to_host(d), h)); }
        CHECK(std::strstr(xtb_last_kernel(), "interp") == nullptr);
        { xtb::xarray<int> d = xt::cumsum(dc, 0); xt::xarray<int> h = xt::cumsum(c, 0); CHECK(same_bits(xtb::to_host(d), h)); }
    }

    if (g_failed) { std::printf("%d check(s) FAILED\n", g_failed); return 1; }
    std::printf("OK test_dropin_reducers\n");
    return 0;
}
