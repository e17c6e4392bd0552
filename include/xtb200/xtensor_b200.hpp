// xtensor_b200.hpp -- the drop-in boundary between xtensor's expression API and
// libxtb200 (include/xtb200.h).  Header-only host C++20; compiled by the user's
// C++ compiler, never by nvcc.
//
//   #include <xtb200/xtensor_b200.hpp>
//   xtb::xtensor<float, 3> a = xtb::to_device(host_a), b = ..., d = ..., c;
//   xt::noalias(c) = xt::sin(a) * b + 2.0f * d;        // one fused kernel on the B200
//   xtb::xtensor<float, 2> s = xt::sum(a, {0});         // axis-reduction kernel
//   auto host_c = xtb::to_host(c);
//
// How it plugs in (all of it is xtensor's *documented* extension mechanism, see
// docs/source/developer/assignment.rst:126-164 and docs/source/external-structures.rst):
//   * a new expression tag, xtb::b200_expression_tag; containers take it as their last
//     template argument (core/xtensor_forward.hpp:50-55, 113-142) and every composite node
//     derives its tag with expression_tag_and (core/xexpression.hpp:333-393), so one device
//     operand makes the whole tree device-tagged;
//   * extension::*_base_impl<tag, ...> specialisations for the node types on the path
//     (the pattern of optional/xoptional.hpp:324-873);
//   * detail::select_xfunction_expression<tag,...> (core/xoperation.hpp:166-183) and
//     temporary_type_from_tag<tag, T> (core/xexpression_traits.hpp:131-141);
//   * xexpression_assigner_base<tag>::assign_data (core/xassign.hpp:65-75, 439-478): THE
//     kernel launch point -- the expression tree is lowered to a postfix program plus operand
//     descriptors and handed to xtb_assign / xtb_reduce;
//   * storage: xtb::device_uvector<T>, a uvector-shaped owner of device memory
//     (containers/xstorage.hpp:33-128: uninitialised, resize discards).
// There is no CPU fallback: node types that cannot be lowered (index views, arbitrary
// callables) fail at compile time.
#ifndef XTB200_XTENSOR_B200_HPP
#define XTB200_XTENSOR_B200_HPP

#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iterator>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

#include <xtensor/containers/xarray.hpp>
#include <xtensor/containers/xscalar.hpp>
#include <xtensor/containers/xtensor.hpp>
#include <xtensor/core/xassign.hpp>
#include <xtensor/core/xeval.hpp>
#include <xtensor/core/xfunction.hpp>
#include <xtensor/core/xmath.hpp>
#include <xtensor/core/xnoalias.hpp>
#include <xtensor/core/xoperation.hpp>
#include <xtensor/generators/xgenerator.hpp>
#include <xtensor/misc/xsort.hpp>
#include <xtensor/misc/xmanipulation.hpp>
#include <xtensor/reducers/xaccumulator.hpp>
#include <xtensor/reducers/xnorm.hpp>
#include <xtensor/reducers/xreducer.hpp>
#include <xtensor/views/xbroadcast.hpp>
#include <xtensor/views/xstrided_view.hpp>
#include <xtensor/views/xview.hpp>

#include "../xtb200.h"

namespace xtb
{
    struct b200_expression_tag
    {
    };

    struct b200_empty_base
    {
        using expression_tag = b200_expression_tag;
    };

    // ---------------------------------------------------------------- errors
    // C status -> exception, following XTENSOR_THROW (core/xtensor_config.hpp:25-35)
    inline void check(int status)
    {
        if (status == XTB_OK)
        {
            return;
        }
        const std::string msg = xtb_last_error();
        if (status == XTB_ERR_SHAPE)
        {
            XTENSOR_THROW(xt::broadcast_error, msg.c_str());
        }
        XTENSOR_THROW(std::runtime_error, "xtb200: " + msg);
    }

    inline void sync()
    {
        check(xtb_sync());
    }

    // ---------------------------------------------------------------- dtypes
    template <class T>
    struct dtype_of;
#define XTB_DTYPE(T, V)                    \
    template <>                            \
    struct dtype_of<T>                     \
    {                                      \
        static constexpr int value = V;    \
    };
    XTB_DTYPE(bool, XTB_BOOL)
    XTB_DTYPE(signed char, XTB_I8)
    XTB_DTYPE(unsigned char, XTB_U8)
    XTB_DTYPE(short, XTB_I16)
    XTB_DTYPE(unsigned short, XTB_U16)
    XTB_DTYPE(int, XTB_I32)
    XTB_DTYPE(unsigned int, XTB_U32)
    XTB_DTYPE(long, XTB_I64)
    XTB_DTYPE(unsigned long, XTB_U64)
    XTB_DTYPE(long long, XTB_I64)
    XTB_DTYPE(unsigned long long, XTB_U64)
    XTB_DTYPE(float, XTB_F32)
    XTB_DTYPE(double, XTB_F64)
    XTB_DTYPE(char, (std::is_signed<char>::value ? XTB_I8 : XTB_U8))
#undef XTB_DTYPE

    template <class T>
    inline constexpr int dtype_v = dtype_of<std::remove_cv_t<T>>::value;

    constexpr int regtype(int dt)
    {
        return dt < XTB_I32 ? int(XTB_I32) : dt;
    }

    // ---------------------------------------------------------------- device iterators
    // Iterator over device memory: position arithmetic and comparison only.  Dereferencing is deleted, so host
    // code that would walk a device container element by element (a(i, j), fill, nested initializer lists,
    // std::copy into it, a host stepper loop) fails to COMPILE instead of faulting on a device address.
    template <class T>
    class device_iterator
    {
    public:

        using iterator_category = std::random_access_iterator_tag;
        using value_type = std::remove_cv_t<T>;
        using difference_type = std::ptrdiff_t;
        using pointer = T*;
        using reference = T&;

        constexpr device_iterator() noexcept = default;
        constexpr explicit device_iterator(T* p) noexcept : m_p(p) {}
        template <class U, std::enable_if_t<std::is_same<const U, T>::value, int> = 0>
        constexpr device_iterator(const device_iterator<U>& rhs) noexcept : m_p(rhs.get()) {}

        constexpr T* get() const noexcept { return m_p; }

        reference operator*() const = delete;      // device memory is not addressable from the host
        pointer operator->() const = delete;
        reference operator[](difference_type) const = delete;

        constexpr device_iterator& operator++() noexcept { ++m_p; return *this; }
        constexpr device_iterator operator++(int) noexcept { device_iterator t(*this); ++m_p; return t; }
        constexpr device_iterator& operator--() noexcept { --m_p; return *this; }
        constexpr device_iterator operator--(int) noexcept { device_iterator t(*this); --m_p; return t; }
        constexpr device_iterator& operator+=(difference_type n) noexcept { m_p += n; return *this; }
        constexpr device_iterator& operator-=(difference_type n) noexcept { m_p -= n; return *this; }
        friend constexpr device_iterator operator+(device_iterator a, difference_type n) noexcept { return device_iterator(a.m_p + n); }
        friend constexpr device_iterator operator+(difference_type n, device_iterator a) noexcept { return device_iterator(a.m_p + n); }
        friend constexpr device_iterator operator-(device_iterator a, difference_type n) noexcept { return device_iterator(a.m_p - n); }
        friend constexpr difference_type operator-(device_iterator a, device_iterator b) noexcept { return a.m_p - b.m_p; }
        friend constexpr bool operator==(device_iterator a, device_iterator b) noexcept { return a.m_p == b.m_p; }
        friend constexpr bool operator!=(device_iterator a, device_iterator b) noexcept { return a.m_p != b.m_p; }
        friend constexpr bool operator<(device_iterator a, device_iterator b) noexcept { return a.m_p < b.m_p; }
        friend constexpr bool operator>(device_iterator a, device_iterator b) noexcept { return a.m_p > b.m_p; }
        friend constexpr bool operator<=(device_iterator a, device_iterator b) noexcept { return a.m_p <= b.m_p; }
        friend constexpr bool operator>=(device_iterator a, device_iterator b) noexcept { return a.m_p >= b.m_p; }

    private:

        T* m_p = nullptr;
    };

    // ---------------------------------------------------------------- storage
    // uvector-shaped owner of device memory.  Contents are uninitialised; resize discards
    // (uvector::resize_impl, containers/xstorage.hpp:217-228); copy = device-to-device copy.
    // Iterators are device_iterator<T>: positions only, dereferencing does not compile (use xtb::to_host).
    template <class T>
    class device_uvector
    {
    public:

        using allocator_type = std::allocator<T>;
        using value_type = T;
        using reference = T&;
        using const_reference = const T&;
        using pointer = T*;
        using const_pointer = const T*;
        using size_type = std::size_t;
        using difference_type = std::ptrdiff_t;
        using iterator = device_iterator<T>;
        using const_iterator = device_iterator<const T>;
        using reverse_iterator = std::reverse_iterator<iterator>;
        using const_reverse_iterator = std::reverse_iterator<const_iterator>;

        device_uvector() noexcept = default;

        explicit device_uvector(size_type n, const allocator_type& = allocator_type())
        {
            allocate(n);
        }

        // n copies of v (xtensor's default constructors build a 0- or 1-element storage this way)
        device_uvector(size_type n, const_reference v, const allocator_type& = allocator_type())
        {
            allocate(n);
            if (n)
            {
                std::vector<T> host(n, v);
                check(xtb_memcpy(m_ptr, host.data(), n * sizeof(T), XTB_H2D));
                check(xtb_sync());
            }
        }

        device_uvector(const device_uvector& rhs)
        {
            allocate(rhs.m_size);
            if (m_size)
            {
                check(xtb_memcpy(m_ptr, rhs.m_ptr, m_size * sizeof(T), XTB_D2D));
            }
        }

        device_uvector(device_uvector&& rhs) noexcept
            : m_ptr(rhs.m_ptr)
            , m_size(rhs.m_size)
        {
            rhs.m_ptr = nullptr;
            rhs.m_size = 0;
        }

        device_uvector& operator=(const device_uvector& rhs)
        {
            if (this != &rhs)
            {
                resize(rhs.m_size);
                if (m_size)
                {
                    check(xtb_memcpy(m_ptr, rhs.m_ptr, m_size * sizeof(T), XTB_D2D));
                }
            }
            return *this;
        }

        device_uvector& operator=(device_uvector&& rhs) noexcept
        {
            if (this != &rhs)
            {
                release();
                m_ptr = rhs.m_ptr;
                m_size = rhs.m_size;
                rhs.m_ptr = nullptr;
                rhs.m_size = 0;
            }
            return *this;
        }

        ~device_uvector()
        {
            release();
        }

        allocator_type get_allocator() const noexcept { return allocator_type(); }
        bool empty() const noexcept { return m_size == 0; }
        size_type size() const noexcept { return m_size; }

        void resize(size_type n)
        {
            if (n != m_size)
            {
                release();
                allocate(n);
            }
        }

        pointer data() noexcept { return m_ptr; }
        const_pointer data() const noexcept { return m_ptr; }
        iterator begin() noexcept { return iterator(m_ptr); }
        iterator end() noexcept { return iterator(m_ptr + m_size); }
        const_iterator begin() const noexcept { return const_iterator(m_ptr); }
        const_iterator end() const noexcept { return const_iterator(m_ptr + m_size); }
        const_iterator cbegin() const noexcept { return const_iterator(m_ptr); }
        const_iterator cend() const noexcept { return const_iterator(m_ptr + m_size); }
        reverse_iterator rbegin() noexcept { return reverse_iterator(end()); }
        reverse_iterator rend() noexcept { return reverse_iterator(begin()); }
        const_reverse_iterator rbegin() const noexcept { return const_reverse_iterator(end()); }
        const_reverse_iterator rend() const noexcept { return const_reverse_iterator(begin()); }
        const_reverse_iterator crbegin() const noexcept { return rbegin(); }
        const_reverse_iterator crend() const noexcept { return rend(); }

        // element access from the host is a compile-time error (use xtb::to_host)
        reference operator[](size_type i) = delete;
        const_reference operator[](size_type i) const = delete;
        reference front() = delete;
        const_reference front() const = delete;
        reference back() = delete;
        const_reference back() const = delete;

        void swap(device_uvector& rhs) noexcept
        {
            std::swap(m_ptr, rhs.m_ptr);
            std::swap(m_size, rhs.m_size);
        }

    private:

        void allocate(size_type n)
        {
            m_size = n;
            m_ptr = nullptr;
            if (n)
            {
                void* p = nullptr;
                check(xtb_malloc(n * sizeof(T), &p));
                m_ptr = static_cast<pointer>(p);
            }
        }

        void release() noexcept
        {
            if (m_ptr)
            {
                xtb_free(m_ptr);
            }
            m_ptr = nullptr;
            m_size = 0;
        }

        pointer m_ptr = nullptr;
        size_type m_size = 0;
    };

    template <class T>
    inline void swap(device_uvector<T>& a, device_uvector<T>& b) noexcept
    {
        a.swap(b);
    }

    // ---------------------------------------------------------------- foreign device memory
    // device_span<T>: the storage of an adaptor over device memory that somebody else owns (cudaMalloc, a
    // torch tensor's data_ptr(), a DLPack capsule).  Counterpart of xbuffer_adaptor<T*, no_ownership>
    // (containers/xbuffer_adaptor.hpp:365, containers/xadapt.hpp:105-215): never allocates or frees,
    // `resize` to another size is an error, assignment from a temporary copies into the adapted buffer
    // (xbuffer_adaptor::operator=(temporary_type&&), :1058-1064) with a device-to-device copy.
    template <class T>
    class device_span
    {
    public:

        using allocator_type = std::allocator<T>;
        using value_type = T;
        using reference = T&;
        using const_reference = const T&;
        using pointer = T*;
        using const_pointer = const T*;
        using size_type = std::size_t;
        using difference_type = std::ptrdiff_t;
        using iterator = device_iterator<T>;
        using const_iterator = device_iterator<const T>;
        using reverse_iterator = std::reverse_iterator<iterator>;
        using const_reverse_iterator = std::reverse_iterator<const_iterator>;
        using temporary_type = device_uvector<T>;

        device_span() noexcept = default;

        device_span(pointer device_ptr, size_type n) noexcept
            : m_ptr(device_ptr)
            , m_size(n)
        {
        }

        device_span(const device_span&) noexcept = default;
        device_span& operator=(const device_span&) noexcept = default;

        device_span& operator=(temporary_type&& tmp)
        {
            resize(tmp.size());
            if (m_size)
            {
                check(xtb_memcpy(m_ptr, tmp.data(), m_size * sizeof(T), XTB_D2D));
            }
            return *this;
        }

        void resize(size_type n)
        {
            if (n != m_size)
            {
                XTENSOR_THROW(std::runtime_error, "xtb200: an adaptor over foreign device memory cannot be resized");
            }
        }

        bool empty() const noexcept { return m_size == 0; }
        size_type size() const noexcept { return m_size; }
        pointer data() noexcept { return m_ptr; }
        const_pointer data() const noexcept { return m_ptr; }
        iterator begin() noexcept { return iterator(m_ptr); }
        iterator end() noexcept { return iterator(m_ptr + m_size); }
        const_iterator begin() const noexcept { return const_iterator(m_ptr); }
        const_iterator end() const noexcept { return const_iterator(m_ptr + m_size); }
        const_iterator cbegin() const noexcept { return const_iterator(m_ptr); }
        const_iterator cend() const noexcept { return const_iterator(m_ptr + m_size); }
        reverse_iterator rbegin() noexcept { return reverse_iterator(end()); }
        reverse_iterator rend() noexcept { return reverse_iterator(begin()); }
        const_reverse_iterator rbegin() const noexcept { return const_reverse_iterator(end()); }
        const_reverse_iterator rend() const noexcept { return const_reverse_iterator(begin()); }
        const_reverse_iterator crbegin() const noexcept { return const_reverse_iterator(end()); }
        const_reverse_iterator crend() const noexcept { return const_reverse_iterator(begin()); }
        // element access from the host is a compile-time error (use xtb::to_host)
        reference operator[](size_type i) = delete;
        const_reference operator[](size_type i) const = delete;
        reference front() = delete;
        const_reference front() const = delete;
        reference back() = delete;
        const_reference back() const = delete;

        void swap(device_span& rhs) noexcept
        {
            std::swap(m_ptr, rhs.m_ptr);
            std::swap(m_size, rhs.m_size);
        }

    private:

        pointer m_ptr = nullptr;
        size_type m_size = 0;
    };

    template <class T>
    inline void swap(device_span<T>& a, device_span<T>& b) noexcept
    {
        a.swap(b);
    }

    // ---------------------------------------------------------------- containers
    template <class T, std::size_t N, xt::layout_type L = XTENSOR_DEFAULT_LAYOUT>
    using xtensor = xt::xtensor_container<device_uvector<T>, N, L, b200_expression_tag>;

    template <class T, xt::layout_type L = XTENSOR_DEFAULT_LAYOUT>
    using xarray = xt::xarray_container<device_uvector<T>, L, xt::dynamic_shape<std::size_t>, b200_expression_tag>;

    // adaptors over foreign device memory (xarray_adaptor / xtensor_adaptor with the device tag)
    template <class T, xt::layout_type L = XTENSOR_DEFAULT_LAYOUT>
    using xarray_adaptor = xt::xarray_adaptor<device_span<T>, L, xt::dynamic_shape<std::size_t>, b200_expression_tag>;

    template <class T, std::size_t N, xt::layout_type L = XTENSOR_DEFAULT_LAYOUT>
    using xtensor_adaptor = xt::xtensor_adaptor<device_span<T>, N, L, b200_expression_tag>;

    template <class E>
    struct is_b200_expression : std::is_same<xt::xexpression_tag_t<E>, b200_expression_tag>
    {
    };

    // xtb::adapt(device_ptr, shape[, strides]) -- xt::adapt(ptr, size, xt::no_ownership(), shape[, strides])
    // (containers/xadapt.hpp:105-215) for device memory: no copy, no ownership; usable on both sides of an
    // assignment (`xt::noalias(xtb::adapt(out_ptr, shape)) = xt::sin(xtb::adapt(in_ptr, shape))`).
    template <class T, class SC>
    inline xarray_adaptor<T> adapt(T* device_ptr, const SC& shape)
    {
        xt::dynamic_shape<std::size_t> sh(shape.begin(), shape.end());
        std::size_t n = 1;
        for (auto e : sh)
        {
            n *= e;
        }
        return xarray_adaptor<T>(device_span<T>(device_ptr, n), sh);
    }

    // custom strides need a dynamic-layout adaptor, as in the reference (containers/xadapt.hpp:158-180)
    template <class T, class SC, class SS>
    inline xarray_adaptor<T, xt::layout_type::dynamic> adapt(T* device_ptr, const SC& shape, const SS& strides)
    {
        xt::dynamic_shape<std::size_t> sh(shape.begin(), shape.end());
        xt::get_strides_t<xt::dynamic_shape<std::size_t>> st(strides.begin(), strides.end());
        std::size_t span = 1;   // elements covered by the strided view
        std::size_t d = 0;
        for (auto e : sh)
        {
            if (e == 0)
            {
                span = 0;
                break;
            }
            span += (e - 1) * static_cast<std::size_t>(st[d] < 0 ? -st[d] : st[d]);
            ++d;
        }
        return xarray_adaptor<T, xt::layout_type::dynamic>(device_span<T>(device_ptr, span), sh, st);
    }

    template <class T>
    inline xarray_adaptor<T> adapt(T* device_ptr, std::initializer_list<std::size_t> shape)
    {
        return adapt(device_ptr, std::vector<std::size_t>(shape));
    }
}

// ======================================================================== xt:: hooks
namespace xt
{
    namespace extension
    {
        // (b200, xtensor) -> b200 is covered by the generic expression_tag_and rules
        // (core/xexpression.hpp:354-363); mixing with the optional tag is not supported.
        template <class EC, std::size_t N, layout_type L>
        struct xtensor_container_base<EC, N, L, xtb::b200_expression_tag>
        {
            using type = xtb::b200_empty_base;
        };

        template <class EC, layout_type L, class SC>
        struct xarray_container_base<EC, L, SC, xtb::b200_expression_tag>
        {
            using type = xtb::b200_empty_base;
        };

        template <class EC, layout_type L, class SC>
        struct xarray_adaptor_base<EC, L, SC, xtb::b200_expression_tag>
        {
            using type = xtb::b200_empty_base;
        };

        template <class EC, std::size_t N, layout_type L>
        struct xtensor_adaptor_base<EC, N, L, xtb::b200_expression_tag>
        {
            using type = xtb::b200_empty_base;
        };

        template <class F, class... CT>
        struct xfunction_base_impl<xtb::b200_expression_tag, F, CT...>
        {
            using type = xtb::b200_empty_base;
        };

        template <class CT, class X>
        struct xbroadcast_base_impl<xtb::b200_expression_tag, CT, X>
        {
            using type = xtb::b200_empty_base;
        };

        template <class CT, class... S>
        struct xview_base_impl<xtb::b200_expression_tag, CT, S...>
        {
            using type = xtb::b200_empty_base;
        };

        template <class CT, class S, layout_type L, class FST>
        struct xstrided_view_base_impl<xtb::b200_expression_tag, CT, S, L, FST>
        {
            using type = xtb::b200_empty_base;
        };

        template <class F, class CT, class X, class O>
        struct xreducer_base_impl<xtb::b200_expression_tag, F, CT, X, O>
        {
            using type = xtb::b200_empty_base;
        };
    }

    namespace extension
    {
        // xshared_expression<E> (core/xexpression.hpp:510-735; what detail::shared_forward wraps rvalue operands of
        // nanmean / nanvar / variance in) publishes no expression_tag of its own: it has its operand's
        template <class E>
        struct get_expression_tag<xshared_expression<E>> : get_expression_tag<E>
        {
        };
    }

    namespace detail
    {
        template <class F, class... E>
        struct select_xfunction_expression<xtb::b200_expression_tag, F, E...>
        {
            using type = xfunction<F, E...>;
        };
    }

    // an adaptor's temporary is an owning device container (containers/xbuffer_adaptor.hpp:584-609)
    template <class T>
    struct temporary_container<xtb::device_span<T>>
    {
        using type = xtb::device_uvector<T>;
    };

    // xt::eval of a VIEW picks the view's own temporary_type typedef, which the reference hard-wires to a host
    // container (views/xview.hpp:304, views/xstrided_view.hpp:80).  For device-tagged views the temporary must
    // live on the device: this constrained specialisation is preferred over the generic one
    // (core/xexpression_traits.hpp:150-154) for them.
    template <class T>
        requires(std::is_same<xexpression_tag_t<std::decay_t<T>>, xtb::b200_expression_tag>::value)
    struct temporary_type<T, void_t<typename std::decay_t<T>::temporary_type>>
    {
        using type = typename temporary_type_from_tag<xtb::b200_expression_tag, T>::type;
    };

    // temporaries of device expressions are device containers of the same rank / value type
    template <class T>
    struct temporary_type_from_tag<xtb::b200_expression_tag, T>
    {
        using I = std::decay_t<T>;
        using shape_type = typename I::shape_type;
        using value_type = typename I::value_type;
        static constexpr layout_type static_layout = XTENSOR_DEFAULT_LAYOUT;
        template <class S>
        struct for_shape
        {
            using type = xtb::xarray<value_type, static_layout>;
        };
        template <class ST, std::size_t N>
        struct for_shape<std::array<ST, N>>
        {
            using type = xtb::xtensor<value_type, N, static_layout>;
        };
        using type = typename for_shape<shape_type>::type;
    };
}

namespace xtb
{
    // ---------------------------------------------------------------- lowering
    // Host-side IR -> postfix program.  Mirrors xtensor_b200/expr.py (the Python test mirror)
    // instruction for instruction, so both emit the canonical encodings that the compile-time
    // device programs match (xtensor_b200/csrc/xtb_static_programs.cuh).
    namespace lower
    {
        struct context
        {
            xtb_program prog{};
            xtb_operand leaves[XTB_MAX_LEAVES]{};
            const void* leaf_id[XTB_MAX_LEAVES]{};        // identity of the node a leaf came from
            std::vector<std::shared_ptr<void>> keepalive;  // temporaries of nested reducers

            void emit(int op, int type, int src = 0, int arg = 0)
            {
                if (prog.n_insns >= XTB_MAX_INSNS)
                {
                    XTENSOR_THROW(std::runtime_error, "xtb200: expression too long for one kernel");
                }
                prog.insns[prog.n_insns++] = xtb_insn{std::uint8_t(op), std::uint8_t(type), std::uint8_t(src), std::uint8_t(arg)};
            }

            template <class T>
            int imm(T v, int rt)
            {
                if (prog.n_imms >= XTB_MAX_IMMS)
                {
                    XTENSOR_THROW(std::runtime_error, "xtb200: too many scalars in one expression");
                }
                std::uint64_t bits = 0;
                switch (rt)
                {
                    case XTB_I32: { std::int32_t x = static_cast<std::int32_t>(v); std::memcpy(&bits, &x, 4); break; }
                    case XTB_U32: { std::uint32_t x = static_cast<std::uint32_t>(v); std::memcpy(&bits, &x, 4); break; }
                    case XTB_I64: { std::int64_t x = static_cast<std::int64_t>(v); std::memcpy(&bits, &x, 8); break; }
                    case XTB_U64: { std::uint64_t x = static_cast<std::uint64_t>(v); std::memcpy(&bits, &x, 8); break; }
                    case XTB_F32: { float x = static_cast<float>(v); std::memcpy(&bits, &x, 4); break; }
                    default: { double x = static_cast<double>(v); std::memcpy(&bits, &x, 8); break; }
                }
                prog.imms[prog.n_imms] = bits;
                return prog.n_imms++;
            }

            int leaf(const void* id, const xtb_operand& op)
            {
                for (int i = 0; i < prog.n_leaves; ++i)
                {
                    if (leaf_id[i] == id && leaves[i].base == op.base && leaves[i].offset == op.offset)
                    {
                        return i;
                    }
                }
                if (prog.n_leaves >= XTB_MAX_LEAVES)
                {
                    XTENSOR_THROW(std::runtime_error, "xtb200: too many tensor operands in one expression");
                }
                leaves[prog.n_leaves] = op;
                leaf_id[prog.n_leaves] = id;
                return prog.n_leaves++;
            }
        };

        // operand descriptor of anything with the strided data interface
        // (data() / data_offset() / shape() / strides(): containers, strided xview, xstrided_view)
        template <class E>
        inline xtb_operand describe(const E& e)
        {
            using value_type = typename E::value_type;
            xtb_operand op{};
            op.base = const_cast<void*>(static_cast<const void*>(e.data()));
            op.offset = static_cast<std::int64_t>(e.data_offset());
            op.dtype = dtype_v<value_type>;
            op.ndim = static_cast<std::int32_t>(e.dimension());
            if (op.ndim > XTB_MAX_DIM)
            {
                XTENSOR_THROW(std::runtime_error, "xtb200: rank > 8 is not supported");
            }
            auto sh = e.shape();
            auto st = e.strides();
            std::size_t d = 0;
            for (auto it = sh.begin(); it != sh.end(); ++it, ++d)
            {
                op.shape[d] = static_cast<std::int64_t>(*it);
            }
            d = 0;
            for (auto it = st.begin(); it != st.end(); ++it, ++d)
            {
                op.stride[d] = static_cast<std::int64_t>(*it);
            }
            return op;
        }

        // ---- functor -> opcode ---------------------------------------------------------
        template <class F>
        struct opcode_of
        {
            static constexpr int value = -1;
        };
#define XTB_OPCODE(FUNCTOR, OP)                 \
    template <>                                 \
    struct opcode_of<FUNCTOR>                   \
    {                                           \
        static constexpr int value = OP;        \
    };
        XTB_OPCODE(xt::detail::negate, XTB_OP_NEG)
        XTB_OPCODE(xt::detail::logical_not, XTB_OP_NOT)
        XTB_OPCODE(xt::detail::bitwise_not, XTB_OP_BITNOT)
        XTB_OPCODE(xt::detail::plus, XTB_OP_ADD)
        XTB_OPCODE(xt::detail::minus, XTB_OP_SUB)
        XTB_OPCODE(xt::detail::multiplies, XTB_OP_MUL)
        XTB_OPCODE(xt::detail::divides, XTB_OP_DIV)
        XTB_OPCODE(xt::detail::modulus, XTB_OP_MOD)
        XTB_OPCODE(xt::detail::logical_or, XTB_OP_LOR)
        XTB_OPCODE(xt::detail::logical_and, XTB_OP_LAND)
        XTB_OPCODE(xt::detail::bitwise_or, XTB_OP_BOR)
        XTB_OPCODE(xt::detail::bitwise_and, XTB_OP_BAND)
        XTB_OPCODE(xt::detail::bitwise_xor, XTB_OP_BXOR)
        XTB_OPCODE(xt::detail::left_shift, XTB_OP_SHL)
        XTB_OPCODE(xt::detail::right_shift, XTB_OP_SHR)
        XTB_OPCODE(xt::detail::less, XTB_OP_LT)
        XTB_OPCODE(xt::detail::less_equal, XTB_OP_LE)
        XTB_OPCODE(xt::detail::greater, XTB_OP_GT)
        XTB_OPCODE(xt::detail::greater_equal, XTB_OP_GE)
        XTB_OPCODE(xt::detail::equal_to, XTB_OP_EQ)
        XTB_OPCODE(xt::detail::not_equal_to, XTB_OP_NE)
        XTB_OPCODE(xt::detail::conditional_ternary, XTB_OP_WHERE)
        XTB_OPCODE(xt::math::abs_fun, XTB_OP_ABS)
        XTB_OPCODE(xt::math::fabs_fun, XTB_OP_ABS)
        XTB_OPCODE(xt::math::exp_fun, XTB_OP_EXP)
        XTB_OPCODE(xt::math::exp2_fun, XTB_OP_EXP2)
        XTB_OPCODE(xt::math::expm1_fun, XTB_OP_EXPM1)
        XTB_OPCODE(xt::math::log_fun, XTB_OP_LOG)
        XTB_OPCODE(xt::math::log10_fun, XTB_OP_LOG10)
        XTB_OPCODE(xt::math::log2_fun, XTB_OP_LOG2)
        XTB_OPCODE(xt::math::log1p_fun, XTB_OP_LOG1P)
        XTB_OPCODE(xt::math::sqrt_fun, XTB_OP_SQRT)
        XTB_OPCODE(xt::math::cbrt_fun, XTB_OP_CBRT)
        XTB_OPCODE(xt::math::sin_fun, XTB_OP_SIN)
        XTB_OPCODE(xt::math::cos_fun, XTB_OP_COS)
        XTB_OPCODE(xt::math::tan_fun, XTB_OP_TAN)
        XTB_OPCODE(xt::math::asin_fun, XTB_OP_ASIN)
        XTB_OPCODE(xt::math::acos_fun, XTB_OP_ACOS)
        XTB_OPCODE(xt::math::atan_fun, XTB_OP_ATAN)
        XTB_OPCODE(xt::math::sinh_fun, XTB_OP_SINH)
        XTB_OPCODE(xt::math::cosh_fun, XTB_OP_COSH)
        XTB_OPCODE(xt::math::tanh_fun, XTB_OP_TANH)
        XTB_OPCODE(xt::math::asinh_fun, XTB_OP_ASINH)
        XTB_OPCODE(xt::math::acosh_fun, XTB_OP_ACOSH)
        XTB_OPCODE(xt::math::atanh_fun, XTB_OP_ATANH)
        XTB_OPCODE(xt::math::erf_fun, XTB_OP_ERF)
        XTB_OPCODE(xt::math::erfc_fun, XTB_OP_ERFC)
        XTB_OPCODE(xt::math::tgamma_fun, XTB_OP_TGAMMA)
        XTB_OPCODE(xt::math::lgamma_fun, XTB_OP_LGAMMA)
        XTB_OPCODE(xt::math::ceil_fun, XTB_OP_CEIL)
        XTB_OPCODE(xt::math::floor_fun, XTB_OP_FLOOR)
        XTB_OPCODE(xt::math::trunc_fun, XTB_OP_TRUNC)
        XTB_OPCODE(xt::math::round_fun, XTB_OP_ROUND)
        XTB_OPCODE(xt::math::nearbyint_fun, XTB_OP_NEARBYINT)
        XTB_OPCODE(xt::math::rint_fun, XTB_OP_RINT)
        XTB_OPCODE(xt::math::isfinite_fun, XTB_OP_ISFINITE)
        XTB_OPCODE(xt::math::isinf_fun, XTB_OP_ISINF)
        XTB_OPCODE(xt::math::isnan_fun, XTB_OP_ISNAN)
        XTB_OPCODE(xt::math::sign_fun, XTB_OP_SIGN)
        XTB_OPCODE(xt::math::deg2rad, XTB_OP_DEG2RAD)
        XTB_OPCODE(xt::math::rad2deg, XTB_OP_RAD2DEG)
        XTB_OPCODE(xt::math::fmod_fun, XTB_OP_FMOD)
        XTB_OPCODE(xt::math::remainder_fun, XTB_OP_REMAINDER)
        XTB_OPCODE(xt::math::fmax_fun, XTB_OP_FMAX)
        XTB_OPCODE(xt::math::fmin_fun, XTB_OP_FMIN)
        XTB_OPCODE(xt::math::fdim_fun, XTB_OP_FDIM)
        XTB_OPCODE(xt::math::pow_fun, XTB_OP_POW)
        XTB_OPCODE(xt::math::hypot_fun, XTB_OP_HYPOT)
        XTB_OPCODE(xt::math::atan2_fun, XTB_OP_ATAN2)
        XTB_OPCODE(xt::math::maximum<void>, XTB_OP_MAXIMUM)
        XTB_OPCODE(xt::math::minimum<void>, XTB_OP_MINIMUM)
        XTB_OPCODE(xt::math::fma_fun, XTB_OP_FMA)
        XTB_OPCODE(xt::math::clamp_fun, XTB_OP_CLAMP)
#undef XTB_OPCODE

        constexpr bool is_compare(int op)
        {
            return (op >= XTB_OP_LT && op <= XTB_OP_NE) || op == XTB_OP_LOR || op == XTB_OP_LAND;
        }

        constexpr bool is_predicate(int op)
        {
            return op == XTB_OP_NOT || op == XTB_OP_ISFINITE || op == XTB_OP_ISINF || op == XTB_OP_ISNAN;
        }

        // lambda_adapt<L> (xt::square, xt::cube; core/xmath.hpp:1034-1127): trace the lambda with a
        // symbolic argument that records multiplications
        struct sym
        {
            int muls;  // number of factors of x
        };

        inline sym operator*(const sym& a, const sym& b)
        {
            return sym{a.muls + b.muls};
        }

        template <class F>
        struct is_lambda_adapt : std::false_type
        {
        };

        template <class L>
        struct is_lambda_adapt<xt::detail::lambda_adapt<L>> : std::true_type
        {
        };

        template <class E>
        struct is_scalar_node : std::false_type
        {
        };

        template <class CT>
        struct is_scalar_node<xt::xscalar<CT>> : std::true_type
        {
        };

        template <class E>
        struct is_function_node : std::false_type
        {
        };

        template <class F, class... CT>
        struct is_function_node<xt::xfunction<F, CT...>> : std::true_type
        {
        };

        template <class E>
        struct is_reducer_node : std::false_type
        {
        };

        template <class F, class CT, class X, class O>
        struct is_reducer_node<xt::xreducer<F, CT, X, O>> : std::true_type
        {
        };

        template <class E>
        struct is_broadcast_node : std::false_type
        {
        };

        template <class CT, class X>
        struct is_broadcast_node<xt::xbroadcast<CT, X>> : std::true_type
        {
        };

        template <class E>
        struct is_generator_node : std::false_type
        {
        };

        template <class F, class R, class S>
        struct is_generator_node<xt::xgenerator<F, R, S>> : std::true_type
        {
        };

        template <class E>
        struct is_strided_view_node : std::false_type
        {
        };

        template <class CT, class S, xt::layout_type L, class FST>
        struct is_strided_view_node<xt::xstrided_view<CT, S, L, FST>> : std::true_type
        {
        };

        // a strided view whose operand is a lazy expression (no data interface anywhere): materialised on use
        template <class E>
        struct is_lazy_view_node : std::false_type
        {
        };

        template <class CT, class S, xt::layout_type L, class FST>
        struct is_lazy_view_node<xt::xstrided_view<CT, S, L, FST>>
            : std::bool_constant<!xt::has_data_interface<xt::xstrided_view<CT, S, L, FST>>::value
                                 && !xt::has_data_interface<std::decay_t<CT>>::value>
        {
        };

        // a node a binary instruction can fetch by itself: a scalar, or a strided leaf whose
        // storage dtype already is the operand register type
        template <class E>
        constexpr bool is_simple(int t)
        {
            using D = std::decay_t<E>;
            if constexpr (is_scalar_node<D>::value)
            {
                return true;
            }
            else if constexpr (is_function_node<D>::value || is_reducer_node<D>::value || is_broadcast_node<D>::value
                               || is_generator_node<D>::value)
            {
                return false;
            }
            else if constexpr (is_lazy_view_node<D>::value)
            {
                return false;   // a view over a lazy expression is materialised first
            }
            else
            {
                return dtype_v<typename D::value_type> == t && t >= XTB_I32;
            }
        }

        template <class E>
        int emit_value(context& c, const E& e, int want);

        // reshape_view(container, shape) (views/xstrided_view.hpp:850-878, used by xt::variance) is an
        // xstrided_view over a *flat adaptor*: it has shape / strides / offset but no data().  Over a
        // contiguous container the flat index is the storage index, so it is still an affine leaf.
        template <class E>
        inline xtb_operand describe_flat_view(const E& e)
        {
            const auto& inner = e.expression();
            using I = std::decay_t<decltype(inner)>;
            static_assert(xt::has_data_interface<I>::value, "xtb200: reshape_view over a non-container expression cannot be lowered");
            if (!inner.is_contiguous())
            {
                XTENSOR_THROW(std::runtime_error, "xtb200: reshape_view needs a contiguous operand");
            }
            using value_type = typename E::value_type;
            xtb_operand op{};
            op.base = const_cast<void*>(static_cast<const void*>(inner.data()));
            op.offset = static_cast<std::int64_t>(inner.data_offset() + e.data_offset());
            op.dtype = dtype_v<value_type>;
            op.ndim = static_cast<std::int32_t>(e.dimension());
            std::size_t d = 0;
            for (auto it = e.shape().begin(); it != e.shape().end(); ++it, ++d)
            {
                op.shape[d] = static_cast<std::int64_t>(*it);
            }
            d = 0;
            for (auto it = e.strides().begin(); it != e.strides().end(); ++it, ++d)
            {
                op.stride[d] = static_cast<std::int64_t>(*it);
            }
            return op;
        }

        template <class E>
        inline xtb_operand describe_leaf(const E& e)
        {
            if constexpr (xt::has_data_interface<E>::value)
            {
                return describe(e);
            }
            else
            {
                static_assert(is_strided_view_node<E>::value,
                              "xtb200: this expression node has no strided data interface and cannot be "
                              "lowered (index / filter / keep-drop views are out of scope; there is no CPU fallback)");
                return describe_flat_view(e);
            }
        }

        template <class E>
        struct is_leaf_node
            : std::bool_constant<!is_scalar_node<E>::value && !is_function_node<E>::value && !is_reducer_node<E>::value
                                 && !is_broadcast_node<E>::value && !is_generator_node<E>::value>
        {
        };

        template <class E>
        std::pair<int, int> fused_src(context& c, const E& e, int t)
        {
            using D = std::decay_t<E>;
            if constexpr (is_scalar_node<D>::value)
            {
                return {XTB_SRC_IMM, c.imm(e(), t)};
            }
            else if constexpr (is_leaf_node<D>::value)
            {
                return {XTB_SRC_LEAF, c.leaf(std::addressof(e), describe_leaf(e))};
            }
            else
            {
                return {0, 0};  // never reached: is_simple() is false for composite nodes
            }
        }

        template <class R, class E>
        xtb::xarray<typename R::value_type> eval_reducer(const R& r);

        template <class F, class... CT>
        int emit_function(context& c, const xt::xfunction<F, CT...>& f)
        {
            using fun_t = xt::xfunction<F, CT...>;
            using value_type = typename fun_t::value_type;
            constexpr std::size_t N = sizeof...(CT);
            const auto& args = f.arguments();
            if constexpr (is_lambda_adapt<F>::value)
            {
                static_assert(N == 1, "only unary lambda functors (square, cube) can be lowered");
                const int t = regtype(dtype_v<value_type>);
                emit_value(c, std::get<0>(args), t);
                // trace: count the factors of x in the lambda's product
                const auto& lam = f.functor();
                const sym r = lam(sym{1});
                if (r.muls == 2) c.emit(XTB_OP_SQUARE, t);
                else if (r.muls == 3) c.emit(XTB_OP_CUBE, t);
                else XTENSOR_THROW(std::runtime_error, "xtb200: unsupported lambda functor");
                return t;
            }
            else if constexpr (opcode_of<F>::value < 0)
            {
                // cast<R>::functor: value_type is R
                static_assert(N == 1, "xtb200: this functor cannot be lowered to a device opcode");
                using arg_t = typename std::decay_t<std::tuple_element_t<0, std::tuple<CT...>>>::value_type;
                const int from = regtype(dtype_v<arg_t>);
                emit_value(c, std::get<0>(args), from);
                c.emit(XTB_OP_CAST, from, 0, dtype_v<value_type>);
                return regtype(dtype_v<value_type>);
            }
            else if constexpr (N == 1)
            {
                constexpr int op = opcode_of<F>::value;
                using arg_t = typename std::decay_t<std::tuple_element_t<0, std::tuple<CT...>>>::value_type;
                // math functors compute in their result type (std::sin(int) is double); predicates
                // and - ~ ! in the promoted operand type
                const int t = is_predicate(op) ? regtype(dtype_v<arg_t>) : regtype(dtype_v<value_type>);
                emit_value(c, std::get<0>(args), t);
                c.emit(op, t);
                return is_predicate(op) ? int(XTB_I32) : t;
            }
            else if constexpr (N == 2)
            {
                constexpr int op = opcode_of<F>::value;
                using A = typename std::decay_t<std::tuple_element_t<0, std::tuple<CT...>>>::value_type;
                using B = typename std::decay_t<std::tuple_element_t<1, std::tuple<CT...>>>::value_type;
                int t;
                if constexpr (op == XTB_OP_SHL || op == XTB_OP_SHR)
                {
                    t = regtype(dtype_v<value_type>);
                }
                else if constexpr (is_compare(op))
                {
                    t = regtype(dtype_v<decltype(std::declval<A>() + std::declval<B>())>);
                }
                else
                {
                    t = regtype(dtype_v<value_type>);
                }
                const auto& a = std::get<0>(args);
                const auto& b = std::get<1>(args);
                using EA = std::decay_t<decltype(a)>;
                using EB = std::decay_t<decltype(b)>;
                if (is_simple<EB>(t))
                {
                    emit_value(c, a, t);
                    auto [src, arg] = fused_src(c, b, t);
                    c.emit(op, t, src, arg);
                }
                else if (is_simple<EA>(t))
                {
                    emit_value(c, b, t);
                    auto [src, arg] = fused_src(c, a, t);
                    c.emit(op, t, src | XTB_SRC_REV, arg);
                }
                else
                {
                    emit_value(c, a, t);
                    emit_value(c, b, t);
                    c.emit(op, t, XTB_SRC_STACK, 0);
                }
                return is_compare(op) ? int(XTB_I32) : t;
            }
            else
            {
                static_assert(N == 3, "xtb200: functors of arity > 3 cannot be lowered");
                constexpr int op = opcode_of<F>::value;
                const int t = regtype(dtype_v<value_type>);
                if constexpr (op == XTB_OP_WHERE)
                {
                    const int ct = emit_value(c, std::get<0>(args), -1);
                    if (ct != XTB_I32)
                    {
                        c.emit(XTB_OP_CAST, ct, 0, XTB_BOOL);
                    }
                }
                else
                {
                    emit_value(c, std::get<0>(args), t);
                }
                emit_value(c, std::get<1>(args), t);
                emit_value(c, std::get<2>(args), t);
                c.emit(op, t);
                return t;
            }
        }

        // leave e's value on the stack in register type `want` (-1: whatever it naturally is)
        template <class E>
        int emit_value(context& c, const E& e, int want)
        {
            using D = std::decay_t<E>;
            int rt;
            if constexpr (is_scalar_node<D>::value)
            {
                using T = typename D::value_type;
                rt = want < 0 ? regtype(dtype_v<T>) : want;
                c.emit(XTB_OP_PUSH, rt, XTB_SRC_IMM, c.imm(e(), rt));
                return rt;
            }
            else if constexpr (is_function_node<D>::value)
            {
                rt = emit_function(c, e);
            }
            else if constexpr (is_broadcast_node<D>::value)
            {
                // explicit broadcast == stride-0 descriptor: the operand broadcasts against the output
                return emit_value(c, e.expression(), want);
            }
            else if constexpr (is_reducer_node<D>::value)
            {
                // nested reducers are materialised first (what xt::eval would do)
                auto tmp = std::make_shared<xtb::xarray<typename D::value_type>>(eval_reducer<D, E>(e));
                c.keepalive.push_back(tmp);
                c.emit(XTB_OP_PUSH, dtype_v<typename D::value_type>, XTB_SRC_LEAF, c.leaf(tmp.get(), describe(*tmp)));
                rt = regtype(dtype_v<typename D::value_type>);
            }
            else if constexpr (is_generator_node<D>::value)
            {
                // generators (arange, linspace, eye, random: generators/xgenerator.hpp, xbuilder.hpp) are host fills:
                // xtensor's own code produces the values once, the result is uploaded and becomes an ordinary leaf
                using T = typename D::value_type;
                xt::xarray<T> host = e;
                auto tmp = std::make_shared<xtb::xarray<T>>();
                std::vector<std::size_t> shp(host.shape().begin(), host.shape().end());
                tmp->resize(shp);
                if (host.size())
                {
                    check(xtb_memcpy(tmp->data(), host.data(), host.size() * sizeof(T), XTB_H2D));
                    check(xtb_sync());   // `host` dies at the end of this scope
                }
                c.keepalive.push_back(tmp);
                c.emit(XTB_OP_PUSH, dtype_v<T>, XTB_SRC_LEAF, c.leaf(tmp.get(), describe(*tmp)));
                rt = regtype(dtype_v<T>);
            }
            else if constexpr (is_lazy_view_node<D>::value)
            {
                // reshape_view / strided_view over a LAZY expression (xt::nanvar reshapes the un-evaluated inner mean,
                // core/xmath.hpp:2785-2808): evaluate the operand into a device temporary (what the flat adaptor would
                // walk element by element), then the view is an affine leaf over that dense buffer
                using T = typename D::value_type;
                auto tmp = std::make_shared<xtb::xarray<T>>();
                *tmp = e.expression();
                c.keepalive.push_back(tmp);
                xtb_operand op{};
                op.base = static_cast<void*>(tmp->data());
                op.offset = static_cast<std::int64_t>(e.data_offset());
                op.dtype = dtype_v<T>;
                op.ndim = static_cast<std::int32_t>(e.dimension());
                std::size_t d = 0;
                for (auto it = e.shape().begin(); it != e.shape().end(); ++it, ++d)
                {
                    op.shape[d] = static_cast<std::int64_t>(*it);
                }
                d = 0;
                for (auto it = e.strides().begin(); it != e.strides().end(); ++it, ++d)
                {
                    op.stride[d] = static_cast<std::int64_t>(*it);
                }
                c.emit(XTB_OP_PUSH, dtype_v<T>, XTB_SRC_LEAF, c.leaf(tmp.get(), op));
                rt = regtype(dtype_v<T>);
            }
            else
            {
                c.emit(XTB_OP_PUSH, dtype_v<typename D::value_type>, XTB_SRC_LEAF, c.leaf(std::addressof(e), describe_leaf(e)));
                rt = regtype(dtype_v<typename D::value_type>);
            }
            if (want >= 0 && rt != want)
            {
                c.emit(XTB_OP_CAST, rt, 0, want);
                rt = want;
            }
            return rt;
        }

        // ---- reducers ---------------------------------------------------------------------
        template <class F>
        struct reduce_op_of
        {
            static constexpr int value = -1;
        };

        template <>
        struct reduce_op_of<xt::detail::plus>
        {
            static constexpr int value = XTB_RED_SUM;
        };

        template <>
        struct reduce_op_of<xt::detail::multiplies>
        {
            static constexpr int value = XTB_RED_PROD;
        };

        template <>
        struct reduce_op_of<xt::math::maximum<void>>
        {
            static constexpr int value = XTB_RED_MAX;
        };

        template <>
        struct reduce_op_of<xt::math::minimum<void>>
        {
            static constexpr int value = XTB_RED_MIN;
        };

        // nan_plus / nan_multiplies ("!isnan(rhs) ? lhs (+|*) rhs : lhs", core/xmath.hpp:2365-2381) are the
        // plain merges over the operand with its NaNs replaced by the merge's identity; the replacement
        // is fused into the reduction kernel's map program (no temporary)
        template <>
        struct reduce_op_of<xt::detail::nan_plus>
        {
            static constexpr int value = XTB_RED_SUM;
        };

        template <>
        struct reduce_op_of<xt::detail::nan_multiplies>
        {
            static constexpr int value = XTB_RED_PROD;
        };

        // nan_min / nan_max (core/xmath.hpp:2333-2363) are native merges of the reduction kernels
        template <>
        struct reduce_op_of<xt::detail::nan_min>
        {
            static constexpr int value = XTB_RED_NANMIN;
        };

        template <>
        struct reduce_op_of<xt::detail::nan_max>
        {
            static constexpr int value = XTB_RED_NANMAX;
        };

        template <class F>
        struct nan_fill_of
        {
            static constexpr int value = -1;   // not a nan-skipping merge
        };

        template <>
        struct nan_fill_of<xt::detail::nan_plus>
        {
            static constexpr int value = 0;
        };

        template <>
        struct nan_fill_of<xt::detail::nan_multiplies>
        {
            static constexpr int value = 1;
        };

        // program + leaves of a reducer's operand, NaNs replaced when the merge skips them
        template <class F, class E>
        inline void emit_reducer_operand(context& c, const E& e)
        {
            constexpr int fill = nan_fill_of<F>::value;
            using vt = typename std::decay_t<E>::value_type;
            if constexpr (fill >= 0 && std::is_floating_point<vt>::value)
            {
                auto mapped = xt::where(xt::isnan(e), vt(fill), e);
                emit_value(c, mapped, -1);
            }
            else
            {
                emit_value(c, e, -1);
            }
        }

        template <class R>
        struct reducer_traits;

        template <class F, class CT, class X, class O>
        struct reducer_traits<xt::xreducer<F, CT, X, O>>
        {
            using options_type = O;
        };

        // xreducer keeps its axes private.  build_reducer() (reducers/xreducer.hpp:1639-1660) rebinds
        // the same axes onto another expression: probing with a lazy broadcast whose extents are
        // distinct primes reveals which dims the reducer removes (or sets to 1 with keep_dims).
        template <class R>
        inline int reducer_axes(const R& r, std::int32_t (&axes)[XTB_MAX_DIM])
        {
            static constexpr std::size_t primes[XTB_MAX_DIM] = {2, 3, 5, 7, 11, 13, 17, 19};
            const std::size_t nd = r.expression().dimension();
            std::vector<std::size_t> probe_shape(primes, primes + nd);
            using probe_value_type = typename std::decay_t<decltype(r.expression())>::value_type;
            auto probe = r.build_reducer(xt::broadcast(probe_value_type(0), probe_shape));
            const auto& rs = probe.shape();
            int na = 0;
            for (std::size_t d = 0; d < nd; ++d)
            {
                bool kept = false;
                for (auto it = rs.begin(); it != rs.end(); ++it)
                {
                    kept = kept || (*it == primes[d]);
                }
                if (!kept)
                {
                    axes[na++] = static_cast<std::int32_t>(d);
                }
            }
            return na;
        }

        // ---- reducers whose reduce functor is an opaque lambda ------------------------------------
        // count_nonzero (core/xmath.hpp:2478-2541), the norms (reducers/xnorm.hpp:369-433, 493-506) and minmax
        // (core/xmath.hpp:2195-2228) build their reduce functor from a lambda; its type says nothing.  Every one
        // of them has the form  f(r, v) = r (+) g(v)  with (+) given by the merge functor (std::plus / detail::plus
        // -> sum, math::maximum -> max) and g one of {v != 0, |v|, v * v, |v|^p}.  The lambda is called on the
        // host with a few probe values to find out which g it is; anything else is rejected (no CPU fallback).
        enum map_kind
        {
            MAP_NONE = 0,
            MAP_COUNT = 1,   // g(v) = (v != 0)                 count_nonzero, norm_l0
            MAP_ABS = 2,     // g(v) = std::abs(v)              norm_l1, norm_linf
            MAP_SQ = 3,      // g(v) = v * v                    norm_sq
            MAP_POWABS = 4   // g(v) = pow(rt(std::abs(v)), p)  norm_lp_to_p
        };

        struct reducer_plan
        {
            int op = -1;          // xtb_reduce_op
            int map = MAP_NONE;
            double p = 0.0;       // MAP_POWABS exponent
            bool zero_init = false;   // merge the reducer's own init (0) once at the end (max-type norms)
        };

        template <class T>
        struct is_array2 : std::false_type
        {
        };

        template <class T>
        struct is_array2<std::array<T, 2>> : std::true_type
        {
            using element_type = T;
        };

        template <class M>
        struct merge_is_plus : std::false_type
        {
        };

        template <>
        struct merge_is_plus<xt::detail::plus> : std::true_type
        {
        };

        template <class T>
        struct merge_is_plus<std::plus<T>> : std::true_type
        {
        };

        template <class M>
        struct merge_is_max : std::false_type
        {
        };

        template <class T>
        struct merge_is_max<xt::math::maximum<T>> : std::true_type
        {
        };

        template <class RF, class MF, class R, class V>
        inline reducer_plan probe_reducer(const RF& f)
        {
            reducer_plan pl;
            if constexpr (std::is_arithmetic<R>::value && std::is_arithmetic<V>::value && std::is_invocable<const RF&, R, V>::value)
            {
                auto g = [&f](double r, double v) { return static_cast<double>(f(static_cast<R>(r), static_cast<V>(v))); };
                const bool is_signed = std::is_signed<V>::value && !std::is_same<V, bool>::value;
                if constexpr (merge_is_plus<MF>::value)
                {
                    const double z = g(7, 0) - 7, a = g(7, 2) - 7, b = g(7, 3) - 7;
                    const double n = is_signed ? g(7, -3) - 7 : b;
                    if (z != 0)
                    {
                        return pl;
                    }
                    pl.op = XTB_RED_SUM;
                    if (std::is_same<V, bool>::value ? (a == 1) : (a == 1 && b == 1 && n == 1))
                    {
                        pl.map = MAP_COUNT;
                    }
                    else if (a == 2 && b == 3 && n == 3)
                    {
                        pl.map = MAP_ABS;
                    }
                    else if (a == 4 && b == 9 && n == 9)
                    {
                        pl.map = MAP_SQ;
                    }
                    else if (a > 0 && b > 0 && n == b)
                    {
                        // |v|^p: p from the value at 2, confirmed at 3 (the probe runs in V's real type: float
                        // operands give p to ~1e-7 only).  norm_lp_to_p's lambda captures p by value and nothing
                        // else (xnorm.hpp:563-566): when the closure is exactly one double that agrees with the
                        // estimate, that double IS p, bit for bit.
                        double pw = std::log2(a);
                        const double tol = std::is_same<V, float>::value ? 1e-5 : 1e-9;
                        if (std::fabs(std::pow(3.0, pw) - b) <= tol * b)
                        {
                            if constexpr (sizeof(RF) == sizeof(double) && std::is_trivially_copyable<RF>::value)
                            {
                                double captured = 0.0;
                                std::memcpy(&captured, &f, sizeof(double));
                                if (std::fabs(captured - pw) <= tol * std::fabs(pw))
                                {
                                    pw = captured;
                                }
                            }
                            pl.map = MAP_POWABS;
                            pl.p = pw;
                        }
                        else
                        {
                            pl.op = -1;
                        }
                    }
                    else
                    {
                        pl.op = -1;
                    }
                }
                else if constexpr (merge_is_max<MF>::value)
                {
                    // norm_linf: std::max<result_type>(r, std::abs(v)); a NaN never replaces r
                    const double n = is_signed ? g(1, -3) : 3;
                    if (g(7, 2) == 7 && g(1, 2) == 2 && n == 3 && g(0, 0) == 0)
                    {
                        pl.op = XTB_RED_NANMAX;
                        pl.map = MAP_ABS;
                        pl.zero_init = true;
                    }
                }
            }
            return pl;
        }

        template <class FS, class V>
        inline reducer_plan plan_reducer(const FS& functors)
        {
            using RF = std::decay_t<typename FS::reduce_functor_type>;
            using MF = std::decay_t<typename FS::merge_functor_type>;
            using IF = std::decay_t<typename FS::init_functor_type>;
            using R = std::decay_t<decltype(std::declval<RF>()(std::declval<IF>()(), std::declval<V>()))>;
            reducer_plan pl;
            if constexpr (reduce_op_of<RF>::value >= 0)
            {
                pl.op = reduce_op_of<RF>::value;
            }
            else
            {
                pl = probe_reducer<RF, MF, R, V>(std::get<0>(functors));
                if (pl.op < 0)
                {
                    XTENSOR_THROW(std::runtime_error,
                                  "xtb200: this reducer's functor cannot be lowered to a device reduction (supported: sum, prod, "
                                  "amax, amin, nan-variants, count_nonzero, minmax and the norms); there is no CPU fallback");
                }
            }
            return pl;
        }

        // program + leaves of a reducer's operand with the plan's map applied; result in the accumulator type R
        template <class RF, class R, class E>
        inline void emit_planned_operand(context& c, const reducer_plan& pl, const E& e)
        {
            using V = typename std::decay_t<E>::value_type;
            if (pl.map == MAP_NONE)
            {
                emit_reducer_operand<RF>(c, e);
                return;
            }
            const int acc = regtype(dtype_v<R>);
            int rt = emit_value(c, e, -1);
            switch (pl.map)
            {
                case MAP_COUNT:
                    c.emit(XTB_OP_NE, rt, XTB_SRC_IMM, c.imm(0, rt));
                    rt = XTB_I32;
                    break;
                case MAP_ABS:
                    c.emit(XTB_OP_ABS, rt);
                    break;
                case MAP_SQ:
                    c.emit(XTB_OP_SQUARE, rt);
                    break;
                default:
                {
                    // norm_lp_to_p(v, p) = pow(rt(std::abs(v)), rt(p)) with rt = real_promote_type_t<V> (xnorm.hpp:185-190)
                    using PT = std::decay_t<decltype(xt::norm_lp_to_p(std::declval<V>(), 0.0))>;
                    const int prt = regtype(dtype_v<PT>);
                    c.emit(XTB_OP_ABS, rt);
                    if (rt != prt)
                    {
                        c.emit(XTB_OP_CAST, rt, 0, prt);
                    }
                    c.emit(XTB_OP_POW, prt, XTB_SRC_IMM, c.imm(pl.p, prt));
                    rt = prt;
                    break;
                }
            }
            if (rt != acc)
            {
                c.emit(XTB_OP_CAST, rt, 0, dtype_v<R>);
            }
        }

        // descriptor of a container of std::array<T, 2> seen as T elements: component `which` of every pair
        template <class T, class OUT>
        inline xtb_operand describe_component(OUT& out, int which)
        {
            xtb_operand op{};
            op.base = static_cast<void*>(out.data());
            op.offset = 2 * static_cast<std::int64_t>(out.data_offset()) + which;
            op.dtype = dtype_v<T>;
            op.ndim = static_cast<std::int32_t>(out.dimension());
            std::size_t d = 0;
            for (auto it = out.shape().begin(); it != out.shape().end(); ++it, ++d)
            {
                op.shape[d] = static_cast<std::int64_t>(*it);
            }
            d = 0;
            for (auto it = out.strides().begin(); it != out.strides().end(); ++it, ++d)
            {
                op.stride[d] = 2 * static_cast<std::int64_t>(*it);
            }
            return op;
        }

        // the one place a reduction is launched from: reduce `e` over axes[na] with the functor triple `functors`
        // into `out`.  IV: type of xt::initial's value (void: none)
        template <class FS, class E, class OUT, class IV>
        void launch_reducer(const FS& functors, const E& e, const std::int32_t* axes, int na, bool keep, const IV* initial_value,
                            OUT& out, bool allreduce, const xtb_finalize* fin = nullptr)
        {
            using RF = std::decay_t<typename FS::reduce_functor_type>;
            using IF = std::decay_t<typename FS::init_functor_type>;
            using V = typename std::decay_t<E>::value_type;
            using R = std::decay_t<decltype(std::declval<RF>()(std::declval<IF>()(), std::declval<V>()))>;
            const auto& sh = e.shape();
            std::int64_t shape[XTB_MAX_DIM] = {0};
            int nd = 0;
            for (auto it = sh.begin(); it != sh.end(); ++it)
            {
                shape[nd++] = static_cast<std::int64_t>(*it);
            }
            auto run = [&](int op, int acc_rt, context& c, const void* initial, const xtb_operand& oop)
            {
                const int st = xtb_reduce_fin(op, acc_rt, &c.prog, c.leaves, nd, shape, na, axes, keep ? 1 : 0, initial, &oop, allreduce ? 1 : 0, fin);
                if (st == XTB_ERR_AXIS)
                {
                    XTENSOR_THROW(std::runtime_error, xtb_last_error());
                }
                check(st);
            };
            if constexpr (is_array2<R>::value)
            {
                // xt::minmax: r[0] = std::min(r[0], v), r[1] = std::max(r[1], v) from {max(), lowest()}; std::min / std::max
                // never take a NaN operand, so the two components are the NaN-skipping extremes merged once with the init
                using T = typename is_array2<R>::element_type;
                const R lo = std::get<0>(functors)(R{T(5), T(5)}, V(3)), hi = std::get<0>(functors)(R{T(5), T(5)}, V(7));
                if (!(lo[0] == T(3) && lo[1] == T(5) && hi[0] == T(5) && hi[1] == T(7)))
                {
                    XTENSOR_THROW(std::runtime_error, "xtb200: only the minmax pair reducer can be lowered (no CPU fallback)");
                }
                const R init = std::get<1>(functors)();
                using W = std::conditional_t<(sizeof(T) < 4), int, T>;   // the accumulator register type
                for (int which = 0; which < 2; ++which)
                {
                    context c;
                    emit_value(c, e, regtype(dtype_v<T>));
                    const W iv = static_cast<W>(init[which]);
                    run(which == 0 ? XTB_RED_NANMIN : XTB_RED_NANMAX, regtype(dtype_v<T>), c, &iv, describe_component<T>(out, which));
                }
            }
            else
            {
                const reducer_plan pl = plan_reducer<FS, V>(functors);
                context c;
                emit_planned_operand<RF, R>(c, pl, e);
                // xt::initial is handed over in the accumulator's REGISTER type (int for the narrow integers)
                using W = std::conditional_t<(sizeof(R) < 4), int, R>;
                W iv{};
                const void* initial = nullptr;
                if constexpr (!std::is_void<IV>::value)
                {
                    if (initial_value)
                    {
                        iv = static_cast<W>(static_cast<R>(*initial_value));
                        initial = &iv;
                    }
                }
                if (pl.zero_init && !initial)
                {
                    iv = W(0);
                    initial = &iv;
                }
                run(pl.op, regtype(dtype_v<R>), c, initial, describe(out));
            }
        }

        // run xreducer<F, CT, X, O> into `out` (any container with the strided data interface)
        template <class R, class OUT>
        void run_reducer(const R& r, OUT& out, bool allreduce = false, const xtb_finalize* fin = nullptr)
        {
            std::int32_t axes[XTB_MAX_DIM] = {0};
            const int na = reducer_axes(r, axes);
            using options_t = typename reducer_traits<R>::options_type;
            constexpr bool keep = typename options_t::keep_dims();
            if constexpr (options_t::has_initial_value)
            {
                const auto iv = r.options().initial_value;
                launch_reducer(r.functors(), r.expression(), axes, na, keep, &iv, out, allreduce, fin);
            }
            else
            {
                launch_reducer(r.functors(), r.expression(), axes, na, keep, static_cast<const void*>(nullptr), out, allreduce, fin);
            }
        }

        // xt::mean / xt::variance / xt::average end in `sum(...) / scalar` (core/xmath.hpp:1827-1852, 2082-2105):
        // xfunction<divides, xreducer<plus...>, xscalar>.  The division is the epilogue of the reduction's last store
        // (xtb_reduce_fin; mean_functor::finalize, reducers/xblockwise_reducer_functors.hpp:146-186) -- one launch
        // fewer and no temporary, same arithmetic: T(sum) / T(divisor) in the node's value type.
        template <class E>
        struct mean_node : std::false_type
        {
        };

        template <class R, class S>
        struct mean_node<xt::xfunction<xt::detail::divides, R, S>>
            : std::bool_constant<is_reducer_node<std::decay_t<R>>::value && is_scalar_node<std::decay_t<S>>::value>
        {
        };

        template <class E, class OUT>
        bool run_mean_node(const E& e, OUT& out, bool allreduce = false)
        {
            using value_type = typename E::value_type;
            using R = std::decay_t<std::tuple_element_t<0, std::decay_t<decltype(e.arguments())>>>;
            using RF = std::decay_t<typename R::reduce_functor_type>;
            if constexpr ((std::is_same<value_type, float>::value || std::is_same<value_type, double>::value)
                          && reduce_op_of<RF>::value == XTB_RED_SUM && std::is_arithmetic<typename R::value_type>::value
                          && std::is_same<typename OUT::value_type, value_type>::value)
            {
                const auto& red = std::get<0>(e.arguments());
                if (!std::equal(red.shape().begin(), red.shape().end(), out.shape().begin(), out.shape().end()))
                {
                    return false;
                }
                xtb_finalize fin{};
                fin.op = XTB_FIN_DIV;
                fin.type = dtype_v<value_type>;
                const value_type div = static_cast<value_type>(std::get<1>(e.arguments())());
                std::memcpy(&fin.imm, &div, sizeof(div));
                run_reducer(red, out, allreduce, &fin);
                return true;
            }
            else
            {
                return false;
            }
        }

        template <class R, class E>
        xtb::xarray<typename R::value_type> eval_reducer(const R& r)
        {
            xtb::xarray<typename R::value_type> out;
            std::vector<std::size_t> shp(r.shape().begin(), r.shape().end());
            out.resize(shp);
            run_reducer(r, out);
            return out;
        }
    }
}

namespace xt
{
    // ---------------------------------------------------------------- the kernel launch point
    // Same signature and `trivial` meaning as the built-in specialisation
    // (core/xassign.hpp:68-75, 439-478); shape inference / resize of e1 already happened in
    // xexpression_assigner<Tag>::assign_xexpression (:480-486, 570-607), which is generic.
    template <>
    class xexpression_assigner_base<xtb::b200_expression_tag>
    {
    public:

        template <class E1, class E2>
        static void assign_data(xexpression<E1>& e1, const xexpression<E2>& e2, bool /*trivial*/)
        {
            E1& lhs = e1.derived_cast();
            const E2& rhs = e2.derived_cast();
            static_assert(std::is_same<xexpression_tag_t<E1>, xtb::b200_expression_tag>::value,
                          "xtb200: the destination of a device expression must be a device container "
                          "(use xtb::to_host to bring results back)");
            if constexpr (xtb::lower::is_reducer_node<E2>::value)
            {
                xtb::lower::run_reducer(rhs, lhs);
            }
            else if constexpr (xtb::lower::mean_node<E2>::value)
            {
                if (!xtb::lower::run_mean_node(rhs, lhs))
                {
                    xtb::lower::context c;
                    xtb::lower::emit_value(c, rhs, -1);
                    xtb_operand out = xtb::lower::describe(lhs);
                    xtb::check(xtb_assign(&c.prog, &out, c.leaves));
                }
            }
            else
            {
                xtb::lower::context c;
                xtb::lower::emit_value(c, rhs, -1);
                xtb_operand out = xtb::lower::describe(lhs);
                xtb::check(xtb_assign(&c.prog, &out, c.leaves));
            }
        }
    };
}

namespace xt
{
    // ---------------------------------------------------------------- eager reducers
    // xt::sum(e, axes, xt::evaluation_strategy::immediate) reaches reduce_immediate through an
    // unqualified call in detail::reduce_impl (reducers/xreducer.hpp:943-953); the generic version
    // walks e.storage() on the host (:289-565).  This more-constrained overload is selected for
    // device-tagged operands and runs the reduction kernel instead.
    template <class F, class E, class X, class O>
        requires xtb::is_b200_expression<std::decay_t<E>>::value
    inline auto reduce_immediate(F&& f, E&& e, X&& axes, O&& raw_options)
    {
        using functors = std::decay_t<F>;
        using reduce_functor_type = typename functors::reduce_functor_type;
        using init_functor_type = typename functors::init_functor_type;
        using expr_value_type = typename std::decay_t<E>::value_type;
        using result_type = std::decay_t<decltype(std::declval<reduce_functor_type>()(
            std::declval<init_functor_type>()(),
            std::declval<expr_value_type>()
        ))>;
        using options_t = reducer_options<result_type, std::decay_t<O>>;
        options_t options(raw_options);

        const std::size_t nd = e.dimension();
        std::int32_t ax[XTB_MAX_DIM] = {0};
        int na = 0;
        for (auto a : axes)
        {
            ax[na++] = static_cast<std::int32_t>(a);
        }
        // same checks and messages as reducers/xreducer.hpp:336-350 (raised by the library)
        constexpr bool keep = typename options_t::keep_dims();
        std::vector<std::size_t> out_shape;
        for (std::size_t d = 0; d < nd; ++d)
        {
            bool reduced = false;
            for (int i = 0; i < na; ++i)
            {
                reduced = reduced || (static_cast<std::size_t>(ax[i]) == d);
            }
            if (!reduced)
            {
                out_shape.push_back(e.shape()[d]);
            }
            else if (keep)
            {
                out_shape.push_back(1);
            }
        }
        xtb::xarray<result_type> result;
        result.resize(out_shape);
        if constexpr (options_t::has_initial_value)
        {
            const auto iv = options.initial_value;
            xtb::lower::launch_reducer(f, e, ax, na, keep, &iv, result, false);
        }
        else
        {
            xtb::lower::launch_reducer(f, e, ax, na, keep, static_cast<const void*>(nullptr), result, false);
        }
        return result;
    }

    // ---------------------------------------------------------------- accumulators
    // xt::cumsum / xt::cumprod call accumulate() unqualified (core/xmath.hpp:2247-2297); the generic
    // accumulator_impl scans res.storage() on the host (reducers/xaccumulator.hpp:215-341).
    namespace detail
    {
        template <class F, class E>
        inline auto b200_accumulate(F&&, E&& e, int axis)
        {
            using functor = std::decay_t<F>;
            using accumulate_functor_type = typename functor::accumulate_functor_type;
            using init_type = typename functor::init_value_type;
            using expr_value_type = typename std::decay_t<E>::value_type;
            using return_type = std::decay_t<decltype(std::declval<accumulate_functor_type>()(
                std::declval<init_type>(),
                std::declval<expr_value_type>()
            ))>;
            constexpr int op = xtb::lower::reduce_op_of<accumulate_functor_type>::value;
            static_assert(op == XTB_RED_SUM || op == XTB_RED_PROD, "xtb200: only cumsum / cumprod can be lowered");
            auto&& src = xt::eval(e);          // containers pass through, expressions become device temporaries
            if (axis >= static_cast<int>(src.dimension()))
            {
                XTENSOR_THROW(std::runtime_error, "Axis larger than expression dimension in accumulator.");
            }
            xtb::xarray<return_type> result;
            std::vector<std::size_t> shp;
            if (axis < 0)
            {
                shp.push_back(src.size());
            }
            else
            {
                shp.assign(src.shape().begin(), src.shape().end());
            }
            result.resize(shp);
            xtb_operand in = xtb::lower::describe(src);
            xtb_operand out = xtb::lower::describe(result);
            xtb::check(xtb_scan(op, xtb::regtype(xtb::dtype_v<return_type>), &in, axis, &out));
            return result;
        }
    }

    template <class F, class E, class EVS = DEFAULT_STRATEGY_ACCUMULATORS, XTL_REQUIRES(is_evaluation_strategy<EVS>)>
        requires xtb::is_b200_expression<std::decay_t<E>>::value
    inline auto accumulate(F&& f, E&& e, EVS = EVS())
    {
        return detail::b200_accumulate(std::forward<F>(f), std::forward<E>(e), -1);
    }

    template <class F, class E, class EVS = DEFAULT_STRATEGY_ACCUMULATORS>
        requires xtb::is_b200_expression<std::decay_t<E>>::value
    inline auto accumulate(F&& f, E&& e, std::ptrdiff_t axis, EVS = EVS())
    {
        const std::size_t ax = normalize_axis(e.dimension(), axis);
        return detail::b200_accumulate(std::forward<F>(f), std::forward<E>(e), static_cast<int>(ax));
    }
}

namespace xt
{
    // ---------------------------------------------------------------- a op= scalar
    // The generic scalar_computed_assign (core/xassign.hpp:525-537) loops over d.storage() on the
    // host.  For the device tag that member template is specialised: `a += 3.1` becomes the
    // in-place kernel  a = static_cast<T>(a + 3.1)  with the same C++ promotion.
    namespace detail
    {
        template <class F>
        struct b200_scalar_op
        {
            static constexpr int value = -1;
        };
#define XTB_SCALAR_OP(F, OP)                  \
    template <>                               \
    struct b200_scalar_op<F>                  \
    {                                         \
        static constexpr int value = OP;      \
    };
        XTB_SCALAR_OP(std::plus<>, XTB_OP_ADD)
        XTB_SCALAR_OP(std::minus<>, XTB_OP_SUB)
        XTB_SCALAR_OP(std::multiplies<>, XTB_OP_MUL)
        XTB_SCALAR_OP(std::divides<>, XTB_OP_DIV)
        XTB_SCALAR_OP(std::modulus<>, XTB_OP_MOD)
        XTB_SCALAR_OP(std::bit_and<>, XTB_OP_BAND)
        XTB_SCALAR_OP(std::bit_or<>, XTB_OP_BOR)
        XTB_SCALAR_OP(std::bit_xor<>, XTB_OP_BXOR)
#undef XTB_SCALAR_OP
    }

    template <>
    template <class E1, class E2, class F>
    inline void xexpression_assigner<xtb::b200_expression_tag>::scalar_computed_assign(xexpression<E1>& e1, const E2& e2, F&&)
    {
        E1& d = e1.derived_cast();
        using T = typename E1::value_type;
        constexpr int op = detail::b200_scalar_op<std::decay_t<F>>::value;
        static_assert(op >= 0, "xtb200: unsupported scalar computed assignment");
        using common_t = decltype(std::declval<T>() + std::declval<E2>());
        const int t = xtb::regtype(xtb::dtype_v<common_t>);
        xtb::lower::context c;
        xtb::lower::emit_value(c, d, t);
        c.emit(op, t, XTB_SRC_IMM, c.imm(e2, t));
        xtb_operand out = xtb::lower::describe_leaf(d);
        xtb::check(xtb_assign(&c.prog, &out, c.leaves));
    }
}

namespace xtb
{
    // ---------------------------------------------------------------- host <-> device
    // Copy a host xtensor / xarray (or any evaluated host expression) to the device.
    template <class E>
    inline auto to_device(const xt::xexpression<E>& host)
    {
        auto&& h = xt::eval(host.derived_cast());
        using H = std::decay_t<decltype(h)>;
        using T = typename H::value_type;
        using result_type = typename xt::temporary_type_from_tag<b200_expression_tag, H>::type;
        result_type d;
        d.resize(h.shape());
        if (h.size())
        {
            check(xtb_memcpy(d.data(), h.data(), h.size() * sizeof(T), XTB_H2D));
            sync();
        }
        return d;
    }

    // Copy a device container back into the matching host container (blocking).
    template <class T, std::size_t N, xt::layout_type L>
    inline xt::xtensor<T, N, L> to_host(const xtensor<T, N, L>& d)
    {
        xt::xtensor<T, N, L> h;
        h.resize(d.shape());
        if (d.size())
        {
            check(xtb_memcpy(h.data(), const_cast<T*>(d.data()), d.size() * sizeof(T), XTB_D2H));
        }
        return h;
    }

    template <class T, xt::layout_type L>
    inline xt::xarray<T, L> to_host(const xarray<T, L>& d)
    {
        xt::xarray<T, L> h;
        std::vector<std::size_t> shp(d.shape().begin(), d.shape().end());
        h.resize(shp);
        if (d.size())
        {
            check(xtb_memcpy(h.data(), const_cast<T*>(d.data()), d.size() * sizeof(T), XTB_D2H));
        }
        return h;
    }

    // In-place copies between an existing host container (pinned memory makes them asynchronous on the device
    // side) and an existing device container of the same size: no allocation, the shapes must already agree.
    template <class D, class H>
    inline void copy_to_device(D& dev, const H& host)
    {
        using T = typename D::value_type;
        static_assert(std::is_same<T, typename H::value_type>::value, "xtb200: copy_to_device needs equal value types");
        if (dev.size() != host.size())
        {
            XTENSOR_THROW(std::runtime_error, "xtb200: copy_to_device: sizes differ");
        }
        if (host.size())
        {
            check(xtb_memcpy(dev.data(), const_cast<T*>(host.data()), host.size() * sizeof(T), XTB_H2D));
        }
    }

    template <class H, class D>
    inline void copy_to_host(H& host, const D& dev)
    {
        using T = typename D::value_type;
        static_assert(std::is_same<T, typename H::value_type>::value, "xtb200: copy_to_host needs equal value types");
        if (dev.size() != host.size())
        {
            XTENSOR_THROW(std::runtime_error, "xtb200: copy_to_host: sizes differ");
        }
        if (dev.size())
        {
            check(xtb_memcpy(host.data(), const_cast<T*>(dev.data()), dev.size() * sizeof(T), XTB_D2H));   // blocks
        }
    }

    // Evaluate any device expression into a new device container and bring it to the host.
    template <class E, std::enable_if_t<!xt::detail::is_container<E>::value, int> = 0>
    inline auto to_host(const xt::xexpression<E>& e)
    {
        typename xt::temporary_type_from_tag<b200_expression_tag, E>::type tmp = e.derived_cast();
        return to_host(tmp);
    }
}

namespace xt
{
    // ---------------------------------------------------------------- argmin / argmax
    // The generic versions (misc/xsort.hpp:1150-1300) eval(e) and then walk the lanes with host iterators; these
    // more-constrained overloads (same template heads, selected for device-tagged operands) run xtb_argreduce.
    // Results are device containers of std::size_t: 0-d for the flat form, rank - 1 otherwise.
    namespace detail
    {
        template <class E>
        inline auto b200_argreduce(const E& e, int op, std::ptrdiff_t axis, bool flat)
        {
            auto&& ed = xt::eval(e);     // containers pass through, expressions become device temporaries
            using ED = std::decay_t<decltype(ed)>;
            xtb::xarray<std::size_t> result;
            std::vector<std::size_t> shp;
            int ax = -1;
            xtb_operand in = xtb::lower::describe(ed);
            if (!flat)
            {
                ax = static_cast<int>(normalize_axis(ed.dimension(), axis));
                for (std::size_t d = 0; d < ed.dimension(); ++d)
                {
                    if (static_cast<int>(d) != ax)
                    {
                        shp.push_back(ed.shape()[d]);
                    }
                }
            }
            result.resize(shp);
            xtb_operand out = xtb::lower::describe(result);
            if (flat && !ed.is_contiguous())
            {
                // the flattened traversal of a strided view: a dense copy first, as eval() of a view would give
                xtb::xarray<typename ED::value_type> dense = ed;
                xtb_operand din = xtb::lower::describe(dense);
                xtb::check(xtb_argreduce(op, &din, -1, &out));
                return result;
            }
            xtb::check(xtb_argreduce(op, &in, ax, &out));
            return result;
        }
    }

    // ---------------------------------------------------------------- nanmean / nanvar on device operands
    // The generic versions (core/xmath.hpp:2694-2842) wrap an rvalue operand in xshared_expression
    // (detail::shared_forward), whose operand cannot be reached from outside, so a shared lazy NODE cannot be
    // lowered.  These more-constrained overloads are the same compositions with a different sharing rule:
    // an lvalue is referenced (as there), an rvalue container is shared (a shared container is still a strided
    // leaf), and an rvalue lazy node is simply copied into both consumers -- it holds its leaves by reference
    // and is evaluated, fused, inside each reduction kernel.
    namespace detail
    {
        template <class T, class E, class F>
        inline auto b200_twice(E&& e, F&& use)
        {
            using D = std::decay_t<E>;
            if constexpr (std::is_lvalue_reference<E>::value)
            {
                return use(e, e);
            }
            else if constexpr (detail::is_container<D>::value)
            {
                auto sh = xt::make_xshared(std::move(e));
                return use(sh, sh);
            }
            else
            {
                D copy(e);
                return use(std::move(copy), std::move(e));
            }
        }
    }

    template <class T = void, class E, class X, class EVS = DEFAULT_STRATEGY_REDUCERS, XTL_REQUIRES(std::negation<is_reducer_options<X>>)>
        requires xtb::is_b200_expression<std::decay_t<E>>::value
    inline auto nanmean(E&& e, X&& axes, EVS es = EVS())
    {
        auto axes_copy = axes;
        using value_type = typename std::conditional_t<std::is_same<T, void>::value, double, T>;
        using sum_type = typename std::conditional_t<
            std::is_same<T, void>::value,
            typename std::common_type_t<typename std::decay_t<E>::value_type, value_type>,
            T>;
        return detail::b200_twice<T>(std::forward<E>(e), [&](auto&& a, auto&& b) {
            return nansum<sum_type>(std::forward<decltype(a)>(a), std::forward<X>(axes), es)
                   / xt::cast<value_type>(count_nonnan(std::forward<decltype(b)>(b), std::move(axes_copy), es));
        });
    }

    template <class T = void, class E, class EVS = DEFAULT_STRATEGY_REDUCERS, XTL_REQUIRES(is_reducer_options<EVS>)>
        requires xtb::is_b200_expression<std::decay_t<E>>::value
    inline auto nanmean(E&& e, EVS es = EVS())
    {
        using value_type = typename std::conditional_t<std::is_same<T, void>::value, double, T>;
        using sum_type = typename std::conditional_t<
            std::is_same<T, void>::value,
            typename std::common_type_t<typename std::decay_t<E>::value_type, value_type>,
            T>;
        return detail::b200_twice<T>(std::forward<E>(e), [&](auto&& a, auto&& b) {
            return nansum<sum_type>(std::forward<decltype(a)>(a), es)
                   / xt::cast<value_type>(count_nonnan(std::forward<decltype(b)>(b), es));
        });
    }

    template <class T = void, class E, class X, class EVS = DEFAULT_STRATEGY_REDUCERS, XTL_REQUIRES(std::negation<is_reducer_options<X>>)>
        requires xtb::is_b200_expression<std::decay_t<E>>::value
    inline auto nanvar(E&& e, X&& axes, EVS es = EVS())
    {
        using result_type = typename std::conditional_t<std::is_same<T, void>::value, double, T>;
        // the inner mean is evaluated once (a small device array), then viewed with keep_dims extents
        auto axes_copy = axes;
        using tmp_shape_t = get_strides_t<typename std::decay_t<E>::shape_type>;
        tmp_shape_t keep_dim_shape = xtl::forward_sequence<tmp_shape_t, decltype(e.shape())>(e.shape());
        for (const auto& el : axes)
        {
            keep_dim_shape[el] = 1;
        }
        return detail::b200_twice<T>(std::forward<E>(e), [&](auto&& a, auto&& b) {
            auto inner_mean = xt::eval(nanmean<result_type>(std::forward<decltype(a)>(a), std::move(axes_copy)));
            auto mrv = reshape_view<XTENSOR_DEFAULT_LAYOUT>(std::move(inner_mean), std::move(keep_dim_shape));
            return nanmean<result_type>(square(cast<result_type>(std::forward<decltype(b)>(b)) - std::move(mrv)), std::forward<X>(axes), es);
        });
    }

    template <class T = void, class E, class EVS = DEFAULT_STRATEGY_REDUCERS, XTL_REQUIRES(is_reducer_options<EVS>)>
        requires xtb::is_b200_expression<std::decay_t<E>>::value
    inline auto nanvar(E&& e, EVS es = EVS())
    {
        return detail::b200_twice<T>(std::forward<E>(e), [&](auto&& a, auto&& b) {
            auto inner_mean = xt::eval(nanmean<T>(std::forward<decltype(a)>(a)));
            return nanmean<T>(square(std::forward<decltype(b)>(b) - std::move(inner_mean)), es);
        });
    }

    // xt::average(e, weights) over the whole array (core/xmath.hpp:1992-2004) reads the weight total on the host
    // (`sum<T>(weights, immediate)()`); for device operands the total stays a 0-d device array and the division
    // is elementwise -- the same arithmetic, sum(e * w) / sum(w)
    template <class T = void, class E, class W, class EVS = DEFAULT_STRATEGY_REDUCERS, XTL_REQUIRES(is_reducer_options<EVS>)>
        requires xtb::is_b200_expression<std::decay_t<E>>::value
    inline auto average(E&& e, W&& weights, EVS ev = EVS())
    {
        if (weights.dimension() != e.dimension()
            || !std::equal(weights.shape().begin(), weights.shape().end(), e.shape().begin()))
        {
            XTENSOR_THROW(std::runtime_error, "Weights need to have the same shape as expression.");
        }
        auto div = sum<T>(weights, evaluation_strategy::immediate);
        return sum<T>(std::forward<E>(e) * std::forward<W>(weights), ev) / std::move(div);
    }

    template <layout_type L = XTENSOR_DEFAULT_TRAVERSAL, class E>
        requires xtb::is_b200_expression<E>::value
    inline auto argmin(const xexpression<E>& e)
    {
        return detail::b200_argreduce(e.derived_cast(), XTB_RED_MIN, 0, true);
    }

    template <layout_type L = XTENSOR_DEFAULT_TRAVERSAL, class E>
        requires xtb::is_b200_expression<E>::value
    inline auto argmin(const xexpression<E>& e, std::ptrdiff_t axis)
    {
        return detail::b200_argreduce(e.derived_cast(), XTB_RED_MIN, axis, false);
    }

    template <layout_type L = XTENSOR_DEFAULT_TRAVERSAL, class E>
        requires xtb::is_b200_expression<E>::value
    inline auto argmax(const xexpression<E>& e)
    {
        return detail::b200_argreduce(e.derived_cast(), XTB_RED_MAX, 0, true);
    }

    template <layout_type L = XTENSOR_DEFAULT_TRAVERSAL, class E>
        requires xtb::is_b200_expression<E>::value
    inline auto argmax(const xexpression<E>& e, std::ptrdiff_t axis)
    {
        return detail::b200_argreduce(e.derived_cast(), XTB_RED_MAX, axis, false);
    }
}

namespace xtb
{
    // ---------------------------------------------------------------- host-resident operands, end to end
    // xtb::assign_host(host_out, expr): evaluate `expr`, whose leaves are HOST containers (pinned memory from
    // xtb::pinned_alloc makes the copies asynchronous), into the host container `host_out` through
    // xtb_assign_host: the leading dimension is cut into chunks and H2D | kernel | D2H of consecutive chunks are
    // pipelined on three streams, so the call costs about max(bytes in, bytes out) / PCIe bandwidth instead of
    // to_device + kernel + to_host back to back.  `expr` is an ordinary xtensor expression over host containers;
    // nothing is evaluated on the CPU.
    template <class OUT, class E>
    inline void assign_host(OUT& host_out, const xt::xexpression<E>& expr, std::int64_t chunk_bytes = 0)
    {
        const E& e = expr.derived_cast();
        std::vector<std::size_t> shp(e.shape().begin(), e.shape().end());
        if (!std::equal(shp.begin(), shp.end(), host_out.shape().begin(), host_out.shape().end()))
        {
            host_out.resize(shp);
        }
        lower::context c;
        lower::emit_value(c, e, -1);
        xtb_operand out = lower::describe(host_out);
        check(xtb_assign_host(&c.prog, &out, c.leaves, chunk_bytes));
    }

    // page-locked host memory for the operands / result of assign_host (xtb_host_alloc)
    template <class T>
    struct pinned_allocator
    {
        using value_type = T;
        pinned_allocator() noexcept = default;
        template <class U> pinned_allocator(const pinned_allocator<U>&) noexcept {}
        T* allocate(std::size_t n)
        {
            void* p = nullptr;
            check(xtb_host_alloc(n * sizeof(T), &p));
            return static_cast<T*>(p);
        }
        void deallocate(T* p, std::size_t) noexcept { xtb_host_free(p); }
        template <class U> bool operator==(const pinned_allocator<U>&) const noexcept { return true; }
        template <class U> bool operator!=(const pinned_allocator<U>&) const noexcept { return false; }
    };

    // host containers in pinned memory: ordinary xt::xtensor / xt::xarray with another allocator
    template <class T, std::size_t N, xt::layout_type L = XTENSOR_DEFAULT_LAYOUT>
    using pinned_xtensor = xt::xtensor_container<xt::uvector<T, pinned_allocator<T>>, N, L>;
    template <class T, xt::layout_type L = XTENSOR_DEFAULT_LAYOUT>
    using pinned_xarray = xt::xarray_container<xt::uvector<T, pinned_allocator<T>>, L>;

    // ---------------------------------------------------------------- multi-GPU (one process per GPU)
    // The reference's own partial -> merge -> finalize contract is xblockwise_reducer
    // (reducers/xblockwise_reducer.hpp:154-185, functors xblockwise_reducer_functors.hpp:45-260).  Here the
    // containers of one process hold a contiguous block of the LEADING axis; elementwise work and reductions
    // over other axes need no exchange, a reduction over axis 0 merges one partial per rank.
    namespace dist
    {
        // Star rendezvous over TCP on the launcher's MASTER_ADDR (rank 0 listens on MASTER_PORT + port_offset):
        // used once, to hand the NCCL id and the peer-memory handles around.  Not on any data path.
        class rendezvous
        {
        public:

            rendezvous(int rank, int world, const std::string& addr, int port);
            ~rendezvous();
            rendezvous(const rendezvous&) = delete;
            // every rank contributes n bytes; all ranks receive the world * n bytes in rank order
            void allgather(const void* mine, std::size_t n, void* all);

        private:

            int m_rank, m_world;
            int m_listen = -1;
            std::vector<int> m_peers;   // rank 0: socket of every other rank; others: [0] = socket to rank 0
        };

        struct communicator
        {
            int rank = 0;
            int world = 1;
            bool peer_memory = false;   // small allreduces run as one kernel over NVLink peer memory
        };

        // Bring up the exchange: NCCL communicator through the C ABI and -- all ranks or none -- the NVLink
        // peer-memory windows (xtb_comm_p2p_*).  Call after xtb_init(local device).
        communicator init(int rank, int world, const std::string& master_addr, int port);
        // RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT as set by torchrun; also selects the device
        communicator init_from_env(int port_offset = 29);
        inline void finalize() { xtb_comm_destroy(); }

        // [begin, end) of a rank's block of the leading axis (remainder rows go to the low ranks)
        inline std::pair<std::size_t, std::size_t> row_block(std::size_t rows, int rank, int world)
        {
            const std::size_t base = rows / world, rem = rows % world;
            const std::size_t b = rank * base + std::min<std::size_t>(rank, rem);
            return {b, b + base + (static_cast<std::size_t>(rank) < rem ? 1 : 0)};
        }

        // xtb::dist::allreduce(xt::sum(local, {0})): evaluate a reducer over the sharded axis and merge the per-rank
        // partials in the same call (fused into the merge kernel over peer memory when attached, NCCL otherwise);
        // every rank gets the identical result.  xt::initial is applied once, after the merge.
        template <class R>
            requires lower::is_reducer_node<R>::value
        inline auto allreduce(const R& reducer)
        {
            xtb::xarray<typename R::value_type> out;
            std::vector<std::size_t> shp(reducer.shape().begin(), reducer.shape().end());
            out.resize(shp);
            lower::run_reducer(reducer, out, true);
            return out;
        }

        // mean over `axes` (which include the sharded axis 0) of a row-sharded expression: merged sum divided by
        // the GLOBAL count, the division being the epilogue of the merge (mean_functor::finalize,
        // xblockwise_reducer_functors.hpp:175-185).  T as in xt::mean<T>.
        template <class T = void, class OUT, class E, class X>
        inline void mean_into(OUT& out, const xt::xexpression<E>& local, const X& axes, std::size_t global_rows)
        {
            const E& e = local.derived_cast();
            using value_type = typename OUT::value_type;
            std::vector<std::size_t> ax(std::begin(axes), std::end(axes));
            auto red = xt::sum<T>(e, ax);
            double count = 1;
            for (auto a : ax)
            {
                count *= static_cast<double>(a == 0 ? global_rows : e.shape()[a]);
            }
            xtb_finalize fin{};
            fin.op = XTB_FIN_DIV;
            fin.type = dtype_v<value_type>;
            const value_type div = static_cast<value_type>(count);
            std::memcpy(&fin.imm, &div, sizeof(div));
            if (!std::equal(red.shape().begin(), red.shape().end(), out.shape().begin(), out.shape().end()))
            {
                std::vector<std::size_t> shp(red.shape().begin(), red.shape().end());
                out.resize(shp);
            }
            const bool crosses = std::find(ax.begin(), ax.end(), std::size_t(0)) != ax.end();
            lower::run_reducer(red, out, crosses, &fin);
        }

        template <class T = void, class E, class X>
        inline auto mean(const xt::xexpression<E>& local, const X& axes, std::size_t global_rows)
        {
            using value_type = std::conditional_t<std::is_same<T, void>::value, double, T>;
            xtb::xarray<value_type> out;
            mean_into<T>(out, local, axes, global_rows);
            return out;
        }

        // two-pass variance (core/xmath.hpp:2082-2105) of a row-sharded expression: the merged mean is broadcast
        // back (every rank holds it), then mean(square(local - mean)) with a second merge.  `m` = the merged mean
        // over the same axes (dist::mean), kept by the caller for re-use (cfg5 needs it for exp(a - m) as well).
        template <class T = void, class OUT, class E, class M, class X>
        inline void variance_into(OUT& out, const xt::xexpression<E>& local, const M& m, const X& axes, std::size_t global_rows)
        {
            const E& e = local.derived_cast();
            using value_type = typename OUT::value_type;
            std::vector<std::size_t> keep(e.shape().begin(), e.shape().end());
            for (auto a : axes)
            {
                keep[static_cast<std::size_t>(a)] = 1;
            }
            auto mrv = xt::reshape_view(m, keep);
            mean_into<value_type>(out, xt::square(e - mrv), axes, global_rows);   // as the reference: no cast (xmath.hpp:2102)
        }

        template <class T = void, class E, class X>
        inline auto variance(const xt::xexpression<E>& local, const X& axes, std::size_t global_rows)
        {
            using value_type = std::conditional_t<std::is_same<T, void>::value, double, T>;
            auto m = mean<T>(local, axes, global_rows);
            xtb::xarray<value_type> out;
            variance_into<T>(out, local, m, axes, global_rows);
            return out;
        }
    }
}

// ---- rendezvous / bring-up (sockets: POSIX) ------------------------------------------------------------
#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <cstdlib>
#include <thread>

namespace xtb
{
    namespace dist
    {
        namespace detail
        {
            inline void send_all(int fd, const void* p, std::size_t n)
            {
                const char* c = static_cast<const char*>(p);
                while (n)
                {
                    const ssize_t k = ::send(fd, c, n, MSG_NOSIGNAL);
                    if (k <= 0)
                    {
                        XTENSOR_THROW(std::runtime_error, "xtb200: rendezvous send failed");
                    }
                    c += k;
                    n -= static_cast<std::size_t>(k);
                }
            }

            inline void recv_all(int fd, void* p, std::size_t n)
            {
                char* c = static_cast<char*>(p);
                while (n)
                {
                    const ssize_t k = ::recv(fd, c, n, 0);
                    if (k <= 0)
                    {
                        XTENSOR_THROW(std::runtime_error, "xtb200: rendezvous receive failed");
                    }
                    c += k;
                    n -= static_cast<std::size_t>(k);
                }
            }
        }

        inline rendezvous::rendezvous(int rank, int world, const std::string& addr, int port)
            : m_rank(rank)
            , m_world(world)
        {
            if (world <= 1)
            {
                return;
            }
            if (rank == 0)
            {
                m_listen = ::socket(AF_INET, SOCK_STREAM, 0);
                int one = 1;
                ::setsockopt(m_listen, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
                sockaddr_in sa{};
                sa.sin_family = AF_INET;
                sa.sin_addr.s_addr = htonl(INADDR_ANY);
                sa.sin_port = htons(static_cast<std::uint16_t>(port));
                if (::bind(m_listen, reinterpret_cast<sockaddr*>(&sa), sizeof(sa)) != 0 || ::listen(m_listen, world) != 0)
                {
                    XTENSOR_THROW(std::runtime_error, "xtb200: rendezvous cannot listen on port " + std::to_string(port));
                }
                m_peers.assign(static_cast<std::size_t>(world), -1);
                for (int i = 1; i < world; ++i)
                {
                    const int fd = ::accept(m_listen, nullptr, nullptr);
                    if (fd < 0)
                    {
                        XTENSOR_THROW(std::runtime_error, "xtb200: rendezvous accept failed");
                    }
                    std::int32_t r = -1;
                    detail::recv_all(fd, &r, sizeof(r));
                    if (r <= 0 || r >= world || m_peers[static_cast<std::size_t>(r)] >= 0)
                    {
                        XTENSOR_THROW(std::runtime_error, "xtb200: rendezvous got a bad rank");
                    }
                    m_peers[static_cast<std::size_t>(r)] = fd;
                }
            }
            else
            {
                addrinfo hints{}, *res = nullptr;
                hints.ai_family = AF_INET;
                hints.ai_socktype = SOCK_STREAM;
                if (::getaddrinfo(addr.c_str(), std::to_string(port).c_str(), &hints, &res) != 0 || !res)
                {
                    XTENSOR_THROW(std::runtime_error, "xtb200: rendezvous cannot resolve " + addr);
                }
                int fd = -1;
                for (int attempt = 0; attempt < 600 && fd < 0; ++attempt)   // rank 0 may not be listening yet
                {
                    fd = ::socket(AF_INET, SOCK_STREAM, 0);
                    if (::connect(fd, res->ai_addr, res->ai_addrlen) != 0)
                    {
                        ::close(fd);
                        fd = -1;
                        std::this_thread::sleep_for(std::chrono::milliseconds(100));
                    }
                }
                ::freeaddrinfo(res);
                if (fd < 0)
                {
                    XTENSOR_THROW(std::runtime_error, "xtb200: rendezvous cannot reach rank 0 at " + addr + ":" + std::to_string(port));
                }
                int one = 1;
                ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
                const std::int32_t r = rank;
                detail::send_all(fd, &r, sizeof(r));
                m_peers.assign(1, fd);
            }
        }

        inline rendezvous::~rendezvous()
        {
            for (int fd : m_peers)
            {
                if (fd >= 0)
                {
                    ::close(fd);
                }
            }
            if (m_listen >= 0)
            {
                ::close(m_listen);
            }
        }

        inline void rendezvous::allgather(const void* mine, std::size_t n, void* all)
        {
            char* out = static_cast<char*>(all);
            if (m_world <= 1)
            {
                std::memcpy(out, mine, n);
                return;
            }
            if (m_rank == 0)
            {
                std::memcpy(out, mine, n);
                for (int r = 1; r < m_world; ++r)
                {
                    detail::recv_all(m_peers[static_cast<std::size_t>(r)], out + static_cast<std::size_t>(r) * n, n);
                }
                for (int r = 1; r < m_world; ++r)
                {
                    detail::send_all(m_peers[static_cast<std::size_t>(r)], out, n * static_cast<std::size_t>(m_world));
                }
            }
            else
            {
                detail::send_all(m_peers[0], mine, n);
                detail::recv_all(m_peers[0], out, n * static_cast<std::size_t>(m_world));
            }
        }

        inline communicator init(int rank, int world, const std::string& master_addr, int port)
        {
            communicator cm;
            cm.rank = rank;
            cm.world = world;
            if (world <= 1)
            {
                return cm;
            }
            rendezvous rv(rank, world, master_addr, port);
            // NCCL id: created on rank 0, everybody takes rank 0's
            std::vector<char> ids(128 * static_cast<std::size_t>(world), 0);
            char mine[128] = {0};
            if (rank == 0)
            {
                check(xtb_comm_unique_id(mine));
            }
            rv.allgather(mine, 128, ids.data());
            check(xtb_comm_init(rank, world, ids.data()));
            // peer-memory windows: every step is collective -- one rank that cannot export or map makes ALL fall back
            if (world <= 8 && std::getenv("XTB_NO_P2P") == nullptr)
            {
                struct slot { char handle[64]; std::int32_t ok; };
                slot me{};
                me.ok = xtb_comm_p2p_handle(me.handle) == XTB_OK;
                std::vector<slot> every(static_cast<std::size_t>(world));
                rv.allgather(&me, sizeof(slot), every.data());
                bool ok = true;
                std::vector<char> handles(64 * static_cast<std::size_t>(world));
                for (int r = 0; r < world; ++r)
                {
                    ok = ok && every[static_cast<std::size_t>(r)].ok;
                    std::memcpy(handles.data() + 64 * r, every[static_cast<std::size_t>(r)].handle, 64);
                }
                std::int32_t attached = ok && xtb_comm_p2p_attach(handles.data(), world) == XTB_OK;
                std::vector<std::int32_t> flags(static_cast<std::size_t>(world));
                rv.allgather(&attached, sizeof(attached), flags.data());   // also the barrier before anyone's next allreduce
                bool all_ok = true;
                for (auto f : flags)
                {
                    all_ok = all_ok && f;
                }
                if (!all_ok)
                {
                    check(xtb_comm_p2p_attach(nullptr, 0));
                }
                cm.peer_memory = all_ok;
            }
            return cm;
        }

        inline communicator init_from_env(int port_offset)
        {
            auto env_int = [](const char* name, int dflt) {
                const char* v = std::getenv(name);
                return v ? std::atoi(v) : dflt;
            };
            const int rank = env_int("RANK", 0), world = env_int("WORLD_SIZE", 1), local = env_int("LOCAL_RANK", rank);
            const char* addr = std::getenv("MASTER_ADDR");
            check(xtb_init(local));
            return init(rank, world, addr ? addr : "127.0.0.1", env_int("MASTER_PORT", 29500) + port_offset);
        }
    }
}

#endif
