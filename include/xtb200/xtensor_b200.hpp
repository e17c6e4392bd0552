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
#include <cstdint>
#include <cstring>
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
#include <xtensor/misc/xmanipulation.hpp>
#include <xtensor/reducers/xaccumulator.hpp>
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

    // ---------------------------------------------------------------- storage
    // uvector-shaped owner of device memory.  Contents are uninitialised; resize discards
    // (uvector::resize_impl, containers/xstorage.hpp:217-228); copy = device-to-device copy.
    // Iterators are raw device pointers: they are only meaningful to libxtb200; host code
    // must not dereference them (use xtb::to_host).
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
        using iterator = pointer;
        using const_iterator = const_pointer;
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
        iterator begin() noexcept { return m_ptr; }
        iterator end() noexcept { return m_ptr + m_size; }
        const_iterator begin() const noexcept { return m_ptr; }
        const_iterator end() const noexcept { return m_ptr + m_size; }
        const_iterator cbegin() const noexcept { return m_ptr; }
        const_iterator cend() const noexcept { return m_ptr + m_size; }
        reverse_iterator rbegin() noexcept { return reverse_iterator(end()); }
        reverse_iterator rend() noexcept { return reverse_iterator(begin()); }
        const_reverse_iterator rbegin() const noexcept { return const_reverse_iterator(end()); }
        const_reverse_iterator rend() const noexcept { return const_reverse_iterator(begin()); }
        const_reverse_iterator crbegin() const noexcept { return rbegin(); }
        const_reverse_iterator crend() const noexcept { return rend(); }

        // device addresses -- never dereference on the host
        reference operator[](size_type i) { return m_ptr[i]; }
        const_reference operator[](size_type i) const { return m_ptr[i]; }
        reference front() { return m_ptr[0]; }
        const_reference front() const { return m_ptr[0]; }
        reference back() { return m_ptr[m_size - 1]; }
        const_reference back() const { return m_ptr[m_size - 1]; }

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
        using iterator = pointer;
        using const_iterator = const_pointer;
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
        // device addresses: valid for pointer arithmetic only (same contract as device_uvector)
        iterator begin() noexcept { return m_ptr; }
        iterator end() noexcept { return m_ptr + m_size; }
        const_iterator begin() const noexcept { return m_ptr; }
        const_iterator end() const noexcept { return m_ptr + m_size; }
        const_iterator cbegin() const noexcept { return m_ptr; }
        const_iterator cend() const noexcept { return m_ptr + m_size; }
        reverse_iterator rbegin() noexcept { return reverse_iterator(end()); }
        reverse_iterator rend() noexcept { return reverse_iterator(begin()); }
        const_reverse_iterator rbegin() const noexcept { return const_reverse_iterator(end()); }
        const_reverse_iterator rend() const noexcept { return const_reverse_iterator(begin()); }
        const_reverse_iterator crbegin() const noexcept { return const_reverse_iterator(end()); }
        const_reverse_iterator crend() const noexcept { return const_reverse_iterator(begin()); }
        reference operator[](size_type i) { return m_ptr[i]; }
        const_reference operator[](size_type i) const { return m_ptr[i]; }
        reference front() { return m_ptr[0]; }
        const_reference front() const { return m_ptr[0]; }
        reference back() { return m_ptr[m_size - 1]; }
        const_reference back() const { return m_ptr[m_size - 1]; }

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
            else if constexpr (is_function_node<D>::value || is_reducer_node<D>::value || is_broadcast_node<D>::value)
            {
                return false;
            }
            else
            {
                return dtype_v<typename D::value_type> == t && t >= XTB_I32;
            }
        }

        template <class E>
        int emit_value(context& c, const E& e, int want);

        template <class E>
        struct is_strided_view_node : std::false_type
        {
        };

        template <class CT, class S, xt::layout_type L, class FST>
        struct is_strided_view_node<xt::xstrided_view<CT, S, L, FST>> : std::true_type
        {
        };

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
                                 && !is_broadcast_node<E>::value>
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
            auto probe = r.build_reducer(xt::broadcast(char(0), probe_shape));
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

        // run xreducer<F, CT, X, O> into `out` (any container with the strided data interface)
        template <class R, class OUT>
        void run_reducer(const R& r, OUT& out, bool allreduce = false)
        {
            using functors = typename R::reduce_functor_type;
            constexpr int op = reduce_op_of<std::decay_t<functors>>::value;
            static_assert(op >= 0, "xtb200: only sum / prod / amax / amin (and nansum / nanprod) reducers can be lowered");
            using acc_t = typename R::value_type;
            context c;
            emit_reducer_operand<std::decay_t<functors>>(c, r.expression());
            const auto& sh = r.expression().shape();
            std::int64_t shape[XTB_MAX_DIM] = {0};
            int nd = 0;
            for (auto it = sh.begin(); it != sh.end(); ++it)
            {
                shape[nd++] = static_cast<std::int64_t>(*it);
            }
            std::int32_t axes[XTB_MAX_DIM] = {0};
            const int na = reducer_axes(r, axes);
            using options_t = typename reducer_traits<R>::options_type;
            constexpr bool keep = typename options_t::keep_dims();
            const void* initial = nullptr;
            acc_t init_v{};
            if constexpr (options_t::has_initial_value)
            {
                init_v = static_cast<acc_t>(r.options().initial_value);
                initial = &init_v;
            }
            xtb_operand oop = describe(out);
            check(xtb_reduce(op, regtype(dtype_v<acc_t>), &c.prog, c.leaves, nd, shape, na, axes, keep ? 1 : 0, initial, &oop,
                             allreduce ? 1 : 0));
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
        constexpr int op = xtb::lower::reduce_op_of<reduce_functor_type>::value;
        static_assert(op >= 0, "xtb200: only sum / prod / amax / amin (and nansum / nanprod) reducers can be lowered");
        (void) f;

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
        xtb::lower::context c;
        xtb::lower::emit_reducer_operand<reduce_functor_type>(c, e);
        std::int64_t shape[XTB_MAX_DIM] = {0};
        for (std::size_t d = 0; d < nd; ++d)
        {
            shape[d] = static_cast<std::int64_t>(e.shape()[d]);
        }
        const void* initial = nullptr;
        result_type init_v{};
        if constexpr (options_t::has_initial_value)
        {
            init_v = static_cast<result_type>(options.initial_value);
            initial = &init_v;
        }
        xtb_operand oop = xtb::lower::describe(result);
        const int st = xtb_reduce(op, xtb::regtype(xtb::dtype_v<result_type>), &c.prog, c.leaves, static_cast<int>(nd), shape, na, ax,
                                  keep ? 1 : 0, initial, &oop, 0);
        if (st == XTB_ERR_AXIS)
        {
            XTENSOR_THROW(std::runtime_error, xtb_last_error());
        }
        xtb::check(st);
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

    // Evaluate any device expression into a new device container and bring it to the host.
    template <class E, std::enable_if_t<!xt::detail::is_container<E>::value, int> = 0>
    inline auto to_host(const xt::xexpression<E>& e)
    {
        typename xt::temporary_type_from_tag<b200_expression_tag, E>::type tmp = e.derived_cast();
        return to_host(tmp);
    }
}

#endif
