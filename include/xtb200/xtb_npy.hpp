// xtb_npy.hpp -- .npy files <-> device containers (host interop for fixtures).
//
// The reference reads and writes .npy through host containers (xt::load_npy / xt::dump_npy,
// include/xtensor/io/xnpy.hpp:740-800).  A device container has no host-side element access, so the two
// functions here are the reference's own parser / writer plus one copy: the file format, dtype checks and
// fortran-order handling stay xtensor's.
//
//   xtb::xarray<float> a = xtb::load_npy<float>("a.npy");      // file -> host temporary -> HBM
//   xtb::dump_npy("out.npy", xt::exp(a));                       // device expression -> evaluated -> host -> file
#pragma once
#include <string>

#include <xtensor/io/xnpy.hpp>

#include "xtensor_b200.hpp"

namespace xtb
{
    template <class T, xt::layout_type L = xt::layout_type::row_major>
    inline xarray<T, L> load_npy(const std::string& filename)
    {
        xt::xarray<T, L> h = xt::load_npy<T, L>(filename);
        return to_device(h);
    }

    template <class E>
    inline void dump_npy(const std::string& filename, const xt::xexpression<E>& device_expr)
    {
        auto&& d = xt::eval(device_expr.derived_cast());      // device temporary (or the container itself)
        xt::dump_npy(filename, to_host(d));
    }
}
