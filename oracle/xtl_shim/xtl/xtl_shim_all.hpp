// xtl_shim_all.hpp -- a minimal STAND-IN for the parts of xtl 0.8.x that xtensor
// 0.27.1's headers use on and around the assignment / reduction path.
//
// xtl (pinned ^0.8.0 by the reference: CMakeLists.txt:43, environment-dev.yml:6) is not
// installed in this image and cannot be fetched, and without it no xtensor header
// compiles.  xtl contributes no arithmetic to the hot path -- only type traits,
// sequence helpers, closure types, a small type-list library, safe integer
// comparisons and select(c, a, b) == c ? a : b.  This file re-implements exactly
// that surface (written from the published interface; it is NOT a copy of xtl), so
// that (a) the real reference can be compiled here as the parity oracle and CPU
// baseline (oracle/ref), and (b) the header-only boundary include/xtb200/*.hpp can
// be compile- and run-tested against the real xtensor headers.
// It is test infrastructure: a maintainer integrating the backend uses the real xtl.
#ifndef XTL_SHIM_ALL_HPP
#define XTL_SHIM_ALL_HPP

#include <algorithm>
#include <array>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <iterator>
#include <limits>
#include <memory>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#define XTL_VERSION_MAJOR 0
#define XTL_VERSION_MINOR 8
#define XTL_VERSION_PATCH 0

namespace xtl
{
    // ------------------------------------------------------------------ traits
    template <class... B> using conjunction = std::conjunction<B...>;
    template <class... B> using disjunction = std::disjunction<B...>;
    template <class B> using negation = std::negation<B>;

#define XTL_REQUIRES_IMPL(...) std::enable_if_t<xtl::conjunction<__VA_ARGS__>::value, int>
#define XTL_REQUIRES(...) XTL_REQUIRES_IMPL(__VA_ARGS__) = 0

    template <class T> struct is_scalar : std::is_arithmetic<T> {};
    template <class T> struct is_arithmetic : std::is_arithmetic<T> {};
    template <class T> struct is_integral : std::is_integral<T> {};
    template <class T> struct is_signed : std::is_signed<T> {};
    template <class T> struct is_floating_point : std::is_floating_point<T> {};

    template <class T> struct is_complex : std::false_type {};
    template <class T> struct is_complex<std::complex<T>> : std::true_type {};
    template <class T> struct is_gen_complex : is_complex<T> {};
    template <class T> struct is_xcomplex : std::false_type {};

    template <class T> struct complex_value_type { using type = T; };
    template <class T> struct complex_value_type<std::complex<T>> { using type = T; };
    template <class T> using complex_value_type_t = typename complex_value_type<T>::type;

    template <class T> struct is_arithmetic<std::complex<T>> : std::true_type {};
    template <class T> struct is_signed<std::complex<T>> : std::true_type {};

    // promote_type: result type of T0 + T1 + ...; bool + bool stays bool
    template <class... T> struct promote_type;
    template <> struct promote_type<> { using type = void; };
    template <class T> struct promote_type<T> { using type = typename promote_type<T, T>::type; };
    template <class T0, class T1> struct promote_type<T0, T1>
    {
        using type = decltype(std::declval<std::decay_t<T0>>() + std::declval<std::decay_t<T1>>());
    };
    template <class T0, class... REST> struct promote_type<T0, REST...>
    {
        using type = decltype(std::declval<std::decay_t<T0>>() + std::declval<typename promote_type<REST...>::type>());
    };
    template <> struct promote_type<bool> { using type = bool; };
    template <class T> struct promote_type<bool, T> { using type = T; };
    template <class... REST> struct promote_type<bool, REST...> { using type = typename promote_type<bool, typename promote_type<REST...>::type>::type; };
    template <> struct promote_type<bool, bool> { using type = bool; };
    template <class... T> using promote_type_t = typename promote_type<T...>::type;

    // big_promote_type: the widest type of the same family (avoids overflow in sums)
    template <class T, class = void> struct big_promote_type_impl { using type = T; };
    template <class T> struct big_promote_type_impl<T, std::enable_if_t<std::is_integral<T>::value && std::is_signed<T>::value>> { using type = long long; };
    template <class T> struct big_promote_type_impl<T, std::enable_if_t<std::is_integral<T>::value && !std::is_signed<T>::value>> { using type = unsigned long long; };
    template <class T> struct big_promote_type_impl<T, std::enable_if_t<std::is_floating_point<T>::value>> { using type = std::conditional_t<(sizeof(T) > sizeof(double)), T, double>; };
    template <class... T> struct big_promote_type { using type = typename big_promote_type_impl<promote_type_t<T...>>::type; };
    template <class... T> using big_promote_type_t = typename big_promote_type<T...>::type;

    // real_promote_type: type of sqrt(T)
    template <class... T> struct real_promote_type
    {
        using base = promote_type_t<T...>;
        using type = std::conditional_t<std::is_integral<base>::value, double, base>;
    };
    template <class... T> using real_promote_type_t = typename real_promote_type<T...>::type;

    template <class... T> struct bool_promote_type
    {
        using base = promote_type_t<T...>;
        using type = std::conditional_t<std::is_arithmetic<base>::value, bool, base>;
    };
    template <class... T> using bool_promote_type_t = typename bool_promote_type<T...>::type;

    // apply_cv: transfer cv + reference qualifiers of T onto U
    template <class T, class U> struct apply_cv
    {
        using nr = std::remove_reference_t<T>;
        using c = std::conditional_t<std::is_const<nr>::value, std::add_const_t<U>, U>;
        using cv = std::conditional_t<std::is_volatile<nr>::value, std::add_volatile_t<c>, c>;
        using type = std::conditional_t<std::is_lvalue_reference<T>::value, std::add_lvalue_reference_t<cv>,
                                        std::conditional_t<std::is_rvalue_reference<T>::value, std::add_rvalue_reference_t<cv>, cv>>;
    };
    template <class T, class U> using apply_cv_t = typename apply_cv<T, U>::type;

    template <class T> struct constify { using type = std::add_const_t<T>; };
    template <class T> struct constify<T&> { using type = std::add_const_t<T>&; };
    template <class T> using constify_t = typename constify<T>::type;

    template <class... C> constexpr bool check_concept() { return conjunction<C...>::value; }

    // ------------------------------------------------------------------ mpl
    namespace mpl
    {
        template <class... T> struct vector {};

        template <bool B, class T, class F> struct eval_if_c { using type = typename T::type; };
        template <class T, class F> struct eval_if_c<false, T, F> { using type = typename F::type; };
        template <class C, class T, class F> struct eval_if : eval_if_c<C::value, T, F> {};
        template <class C, class T, class F> using eval_if_t = typename eval_if<C, T, F>::type;

        template <bool B> using bool_ = std::integral_constant<bool, B>;
        template <class C, class T, class F> struct if_ { using type = std::conditional_t<C::value, T, F>; };
        template <class C, class T, class F> using if_t = typename if_<C, T, F>::type;

        template <class L> struct size;
        template <template <class...> class L, class... T> struct size<L<T...>> : std::integral_constant<std::size_t, sizeof...(T)> {};

        template <class L> struct front;
        template <template <class...> class L, class T, class... R> struct front<L<T, R...>> { using type = T; };
        template <class L> using front_t = typename front<L>::type;

        template <class L> struct back;
        template <template <class...> class L, class T> struct back<L<T>> { using type = T; };
        template <template <class...> class L, class T, class... R> struct back<L<T, R...>> { using type = typename back<L<R...>>::type; };
        template <class L> using back_t = typename back<L>::type;

        template <class L> struct pop_front;
        template <template <class...> class L, class T, class... R> struct pop_front<L<T, R...>> { using type = L<R...>; };
        template <class L> using pop_front_t = typename pop_front<L>::type;

        template <class L, class T> struct push_back;
        template <template <class...> class L, class... U, class T> struct push_back<L<U...>, T> { using type = L<U..., T>; };
        template <class L, class T> using push_back_t = typename push_back<L, T>::type;

        template <class L, class T> struct push_front;
        template <template <class...> class L, class... U, class T> struct push_front<L<U...>, T> { using type = L<T, U...>; };
        template <class L, class T> using push_front_t = typename push_front<L, T>::type;

        template <class L, class V> struct contains;
        template <template <class...> class L, class V> struct contains<L<>, V> : std::false_type {};
        template <template <class...> class L, class T, class... R, class V>
        struct contains<L<T, R...>, V> : std::conditional_t<std::is_same<T, V>::value, std::true_type, contains<L<R...>, V>> {};

        // find_if: index of the first element satisfying Test (size when none does)
        template <template <class> class Test, class L> struct find_if;
        template <template <class> class Test, template <class...> class L> struct find_if<Test, L<>> : std::integral_constant<std::size_t, 0> {};
        template <template <class> class Test, template <class...> class L, class T, class... R>
        struct find_if<Test, L<T, R...>>
            : std::conditional_t<Test<T>::value, std::integral_constant<std::size_t, 0>,
                                 std::integral_constant<std::size_t, 1 + find_if<Test, L<R...>>::value>> {};

        template <class S, class L> struct cast;
        template <template <class...> class F, class... T, template <class...> class L, class... U> struct cast<F<T...>, L<U...>> { using type = L<T...>; };
        template <class S, class L> using cast_t = typename cast<S, L>::type;

        template <bool cond, class TF, class FF> inline decltype(auto) static_if(const TF& tf, const FF& ff)
        {
            if constexpr (cond) return tf(std::identity{});
            else return ff(std::identity{});
        }
    }

    // ------------------------------------------------------------------ functional
    struct identity
    {
        template <class T> constexpr T&& operator()(T&& x) const noexcept { return std::forward<T>(x); }
    };

    // select(cond, a, b): scalar form is cond ? a : b with the common type of a and b
    template <class B, class T1, class T2,
              XTL_REQUIRES(std::is_arithmetic<std::decay_t<B>>)>
    constexpr std::common_type_t<std::decay_t<T1>, std::decay_t<T2>> select(const B& cond, const T1& v1, const T2& v2) noexcept
    {
        return cond ? v1 : v2;
    }

    // ------------------------------------------------------------------ compare
    template <class T1, class T2> constexpr bool cmp_equal(T1 t, T2 u) noexcept
    {
        using UT = std::make_unsigned_t<T1>;
        using UU = std::make_unsigned_t<T2>;
        if constexpr (std::is_signed<T1>::value == std::is_signed<T2>::value) return t == u;
        else if constexpr (std::is_signed<T1>::value) return t < 0 ? false : UT(t) == u;
        else return u < 0 ? false : t == UU(u);
    }
    template <class T1, class T2> constexpr bool cmp_not_equal(T1 t, T2 u) noexcept { return !cmp_equal(t, u); }
    template <class T1, class T2> constexpr bool cmp_less(T1 t, T2 u) noexcept
    {
        using UT = std::make_unsigned_t<T1>;
        using UU = std::make_unsigned_t<T2>;
        if constexpr (std::is_signed<T1>::value == std::is_signed<T2>::value) return t < u;
        else if constexpr (std::is_signed<T1>::value) return t < 0 ? true : UT(t) < u;
        else return u < 0 ? false : t < UU(u);
    }
    template <class T1, class T2> constexpr bool cmp_greater(T1 t, T2 u) noexcept { return cmp_less(u, t); }
    template <class T1, class T2> constexpr bool cmp_less_equal(T1 t, T2 u) noexcept { return !cmp_greater(t, u); }
    template <class T1, class T2> constexpr bool cmp_greater_equal(T1 t, T2 u) noexcept { return !cmp_less(t, u); }

    // ------------------------------------------------------------------ closures
    template <class S> struct closure_type { using underlying_type = std::conditional_t<std::is_const<std::remove_reference_t<S>>::value, const std::decay_t<S>, std::decay_t<S>>;
                                             using type = typename std::conditional<std::is_lvalue_reference<S>::value, underlying_type&, underlying_type>::type; };
    template <class S> using closure_type_t = typename closure_type<S>::type;
    template <class S> struct const_closure_type { using underlying_type = std::decay_t<S>;
                                                   using type = typename std::conditional<std::is_lvalue_reference<S>::value, std::add_const_t<underlying_type>&, underlying_type>::type; };
    template <class S> using const_closure_type_t = typename const_closure_type<S>::type;
    template <class S> struct ptr_closure_type { using underlying_type = std::conditional_t<std::is_const<std::remove_reference_t<S>>::value, const std::decay_t<S>, std::decay_t<S>>;
                                                 using type = std::conditional_t<std::is_lvalue_reference<S>::value, underlying_type*, underlying_type>; };
    template <class S> using ptr_closure_type_t = typename ptr_closure_type<S>::type;

    // xclosure_wrapper: holds a value or a reference, assignable in both cases
    template <class CT> class xclosure_wrapper
    {
    public:
        using self_type = xclosure_wrapper<CT>;
        using closure_type = CT;
        using const_closure_type = std::add_const_t<CT>;
        using value_type = std::decay_t<CT>;
        using reference = std::conditional_t<std::is_const<std::remove_reference_t<CT>>::value, const value_type&, value_type&>;
        using pointer = std::conditional_t<std::is_const<std::remove_reference_t<CT>>::value, const value_type*, value_type*>;

        xclosure_wrapper(value_type&& e) : m_wrappee(std::move(e)) {}
        xclosure_wrapper(reference e) : m_wrappee(e) {}
        xclosure_wrapper(const self_type& rhs) = default;
        xclosure_wrapper(self_type&& rhs) = default;
        self_type& operator=(const self_type& rhs) { deep_copy(rhs.m_wrappee); return *this; }
        self_type& operator=(self_type&& rhs) { deep_copy(rhs.m_wrappee); return *this; }
        template <class T> self_type& operator=(T&& t) { m_wrappee = std::forward<T>(t); return *this; }
        operator closure_type() noexcept { return m_wrappee; }
        operator const_closure_type() const noexcept { return m_wrappee; }
        std::add_lvalue_reference_t<closure_type> get() & noexcept { return m_wrappee; }
        std::add_lvalue_reference_t<std::add_const_t<closure_type>> get() const& noexcept { return m_wrappee; }
        closure_type get() && noexcept { return m_wrappee; }
        pointer operator&() noexcept { return &m_wrappee; }
        bool equal(const self_type& rhs) const { return &m_wrappee == &rhs.m_wrappee; }
        void swap(self_type& rhs) { using std::swap; swap(m_wrappee, rhs.m_wrappee); }

    private:
        template <class T> void deep_copy(const T& v) { m_wrappee = v; }
        CT m_wrappee;
    };
    template <class T> inline decltype(auto) closure(T&& t) { return xclosure_wrapper<closure_type_t<T>>(std::forward<T>(t)); }
    template <class T> inline decltype(auto) const_closure(T&& t) { return xclosure_wrapper<const_closure_type_t<T>>(std::forward<T>(t)); }

    // xclosure_pointer: pointer-like access to a stored value or reference
    template <class CT> class xclosure_pointer
    {
    public:
        using self_type = xclosure_pointer<CT>;
        using closure_type = CT;
        using value_type = std::decay_t<CT>;
        using const_reference = const value_type&;
        using reference = std::conditional_t<std::is_const<std::remove_reference_t<CT>>::value, const_reference, value_type&>;
        using const_pointer = const value_type*;
        using pointer = std::conditional_t<std::is_const<std::remove_reference_t<CT>>::value, const_pointer, value_type*>;

        xclosure_pointer(value_type&& e) : m_wrappee(std::move(e)) {}
        xclosure_pointer(reference e) : m_wrappee(e) {}
        reference operator*() noexcept { return m_wrappee; }
        const_reference operator*() const noexcept { return m_wrappee; }
        pointer operator->() noexcept { return const_cast<pointer>(std::addressof(m_wrappee)); }
        const_pointer operator->() const noexcept { return std::addressof(m_wrappee); }

    private:
        CT m_wrappee;
    };
    template <class T> inline auto closure_pointer(T&& t) { return xclosure_pointer<closure_type_t<T>>(std::forward<T>(t)); }
    template <class T> inline auto const_closure_pointer(T&& t) { return xclosure_pointer<const_closure_type_t<T>>(std::forward<T>(t)); }

    // ------------------------------------------------------------------ sequences
    namespace detail
    {
        template <class S> struct is_std_array : std::false_type {};
        template <class T, std::size_t N> struct is_std_array<std::array<T, N>> : std::true_type {};

        template <class S, class = void> struct has_resize : std::false_type {};
        template <class S> struct has_resize<S, std::void_t<decltype(std::declval<S&>().resize(std::size_t(0)))>> : std::true_type {};

        template <class S> struct sequence_builder
        {
            using value_type = typename S::value_type;
            using size_type = typename S::size_type;
            static S make(size_type size) { return S(size); }
            static S make(size_type size, value_type v) { return S(size, v); }
            static S make(std::initializer_list<value_type> init) { return S(init); }
        };
        template <class T, std::size_t N> struct sequence_builder<std::array<T, N>>
        {
            using S = std::array<T, N>;
            using value_type = T;
            using size_type = std::size_t;
            static S make(size_type) { return S(); }
            static S make(size_type, value_type v) { S s; s.fill(v); return s; }
            static S make(std::initializer_list<value_type> init) { S s; std::copy(init.begin(), init.end(), s.begin()); return s; }
        };
    }
    template <class S> inline S make_sequence(typename S::size_type size) { return detail::sequence_builder<S>::make(size); }
    template <class S> inline S make_sequence(typename S::size_type size, typename S::value_type v) { return detail::sequence_builder<S>::make(size, v); }
    template <class S> inline S make_sequence(std::initializer_list<typename S::value_type> init) { return detail::sequence_builder<S>::make(init); }

    namespace detail
    {
        template <class R, class A, class = void> struct sequence_forwarder_impl
        {
            template <class T> static R forward(const T& r)
            {
                R ret = make_sequence<R>(static_cast<typename R::size_type>(std::distance(std::begin(r), std::end(r))));
                std::copy(std::begin(r), std::end(r), std::begin(ret));
                return ret;
            }
        };
        template <class R, class A> struct sequence_forwarder_impl<R, A, std::enable_if_t<std::is_same<R, std::decay_t<A>>::value>>
        {
            template <class T> static T&& forward(typename std::remove_reference<T>::type& t) noexcept { return static_cast<T&&>(t); }
            template <class T> static T&& forward(typename std::remove_reference<T>::type&& t) noexcept { return static_cast<T&&>(t); }
        };
    }
    // forward_sequence<R, A>(a): perfect-forward when decay_t<A> == R, else copy-convert
    template <class R, class A> inline decltype(auto) forward_sequence(typename std::remove_reference<A>::type& s)
    {
        using forwarder = detail::sequence_forwarder_impl<std::decay_t<R>, A>;
        return forwarder::template forward<A>(s);
    }
    template <class R, class A> inline decltype(auto) forward_sequence(typename std::remove_reference<A>::type&& s)
    {
        using forwarder = detail::sequence_forwarder_impl<std::decay_t<R>, A>;
        static_assert(!std::is_lvalue_reference<A>::value, "Can not forward an rvalue as an lvalue.");
        return forwarder::template forward<A>(std::move(s));
    }

    template <class T> struct sequence_size { static constexpr std::size_t value = 0; };
    template <class T, std::size_t N> struct sequence_size<std::array<T, N>> { static constexpr std::size_t value = N; };
    template <class T, std::size_t N> struct sequence_size<const std::array<T, N>> { static constexpr std::size_t value = N; };

    // ------------------------------------------------------------------ iterator bases (CRTP operator providers)
    template <class I, class T, class D = std::ptrdiff_t, class P = T*, class R = T&>
    class xbidirectional_iterator_base
    {
    public:
        using derived_type = I;
        using value_type = T;
        using reference = R;
        using pointer = P;
        using difference_type = D;
        using iterator_category = std::bidirectional_iterator_tag;
        inline friend derived_type operator++(derived_type& d, int) { derived_type tmp(d); ++d; return tmp; }
        inline friend derived_type operator--(derived_type& d, int) { derived_type tmp(d); --d; return tmp; }
        inline friend bool operator!=(const derived_type& lhs, const derived_type& rhs) { return !(lhs == rhs); }
    };
    template <class I, class T, class D = std::ptrdiff_t, class P = T*, class R = T&>
    class xrandom_access_iterator_base : public xbidirectional_iterator_base<I, T, D, P, R>
    {
    public:
        using derived_type = I;
        using value_type = T;
        using reference = R;
        using pointer = P;
        using difference_type = D;
        using iterator_category = std::random_access_iterator_tag;
        inline reference operator[](difference_type n) const { return *(*static_cast<const derived_type*>(this) + n); }
        inline friend derived_type operator+(const derived_type& it, difference_type n) { derived_type tmp(it); return tmp += n; }
        inline friend derived_type operator+(difference_type n, const derived_type& it) { derived_type tmp(it); return tmp += n; }
        inline friend derived_type operator-(const derived_type& it, difference_type n) { derived_type tmp(it); return tmp -= n; }
        inline friend bool operator<=(const derived_type& lhs, const derived_type& rhs) { return !(rhs < lhs); }
        inline friend bool operator>=(const derived_type& lhs, const derived_type& rhs) { return !(lhs < rhs); }
        inline friend bool operator>(const derived_type& lhs, const derived_type& rhs) { return rhs < lhs; }
    };
    template <class T> using xrandom_access_iterator_base2 = xrandom_access_iterator_base<typename T::iterator_type, typename T::value_type, typename T::difference_type, typename T::pointer, typename T::reference>;
    template <class T> using xbidirectional_iterator_base2 = xbidirectional_iterator_base<typename T::iterator_type, typename T::value_type, typename T::difference_type, typename T::pointer, typename T::reference>;

    // ------------------------------------------------------------------ complex helpers
    template <class E> inline decltype(auto) forward_real(E&& e) { if constexpr (is_complex<std::decay_t<E>>::value) return e.real(); else return std::forward<E>(e); }
    template <class E> inline decltype(auto) forward_imag(E&& e) { if constexpr (is_complex<std::decay_t<E>>::value) return e.imag(); else return std::decay_t<E>(0); }
    template <class M, std::size_t I, class T> inline decltype(auto) forward_offset(T&& v) noexcept
    {
        if constexpr (is_complex<std::decay_t<T>>::value)
        {
            using real_t = typename std::decay_t<T>::value_type;
            return reinterpret_cast<apply_cv_t<T, real_t>*>(&v)[I / sizeof(real_t)];
        }
        else return std::forward<T>(v);
    }

    // ------------------------------------------------------------------ names only (never instantiated on this path)
    template <class T, class A, class BA> class xoptional_vector;
    template <class CT, class CB> class xoptional;
    template <class B, class A = std::allocator<B>> class xdynamic_bitset;
    template <class T, class B> class xmasked_value;
    template <class T> struct is_xoptional : std::false_type {};
    template <class T> struct is_xmasked_value : std::false_type {};
    template <class P> class xproxy_wrapper_impl;

    template <class T> inline T&& value(T&& v) { return std::forward<T>(v); }
    template <class T> inline bool has_value(T&&) { return true; }

    // ------------------------------------------------------------------ xplatform.hpp (io/xnpy.hpp reads the byte order)
    enum class endian { big_endian, little_endian, mixed };
    inline endian endianness()
    {
        const std::uint32_t probe = 0x01020304u;
        const unsigned char* b = reinterpret_cast<const unsigned char*>(&probe);
        return b[0] == 4 ? endian::little_endian : (b[0] == 1 ? endian::big_endian : endian::mixed);
    }
}

#endif
