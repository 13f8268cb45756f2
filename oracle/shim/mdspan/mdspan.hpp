// TEST INFRASTRUCTURE ONLY (oracle build). Not part of the product.
//
// The reference pulls `Kokkos::mdspan` & friends from the kokkos/mdspan project
// via CMake FetchContent (reference CMakeLists.txt:35). That dependency is not
// vendored and there is no network here, so this header supplies the same names
// on top of libcu++'s `cuda::std::mdspan` (shipped with CUDA 12.9).
//
// Names required by the reference: src/neo/container/mdspan.hpp:31-42.
#pragma once

#include <cuda/std/mdspan>

#include <cstddef>
#include <tuple>
#include <type_traits>
#include <utility>

namespace Kokkos {

using ::cuda::std::default_accessor;
using ::cuda::std::dextents;
using ::cuda::std::dynamic_extent;
using ::cuda::std::extents;
using ::cuda::std::full_extent;
using ::cuda::std::full_extent_t;
using ::cuda::std::layout_left;
using ::cuda::std::layout_right;
using ::cuda::std::layout_stride;
using ::cuda::std::mdspan;

namespace oracle_shim {

// The reference slices with `std::tuple{begin, end}` (e.g. overlap_save.hpp:91);
// libcu++'s submdspan wants a pair-like from its own namespace.
template<typename Slice>
struct slice_adaptor
{
    static constexpr auto apply(Slice s) { return s; }
};

template<typename First, typename Last>
struct slice_adaptor<std::tuple<First, Last>>
{
    static constexpr auto apply(std::tuple<First, Last> const& s)
    {
        using pair_t = ::cuda::std::pair<std::size_t, std::size_t>;
        return pair_t{static_cast<std::size_t>(std::get<0>(s)), static_cast<std::size_t>(std::get<1>(s))};
    }
};

}  // namespace oracle_shim

template<typename Elem, typename Ext, typename Layout, typename Acc, typename... Slices>
constexpr auto submdspan(::cuda::std::mdspan<Elem, Ext, Layout, Acc> const& view, Slices... slices)
{
    return ::cuda::std::submdspan(view, oracle_shim::slice_adaptor<Slices>::apply(slices)...);
}

}  // namespace Kokkos
