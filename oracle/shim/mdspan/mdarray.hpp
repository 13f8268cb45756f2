// TEST INFRASTRUCTURE ONLY (oracle build). Not part of the product.
//
// Minimal owning multi-dimensional array with the subset of the
// `Kokkos::Experimental::mdarray` interface the reference touches
// (ctor from extents / index pack, extent(s), size, data, operator(),
// to_mdspan, implicit conversion to mdspan). Storage is a std::vector.
#pragma once

#include "mdspan.hpp"

#include <vector>

namespace Kokkos::Experimental {

template<typename Elem, typename Ext, typename Layout = ::Kokkos::layout_right, typename Storage = std::vector<Elem>>
class mdarray
{
public:
    using element_type   = Elem;
    using value_type     = std::remove_cv_t<Elem>;
    using extents_type   = Ext;
    using layout_type    = Layout;
    using mapping_type   = typename Layout::template mapping<Ext>;
    using index_type     = typename Ext::index_type;
    using size_type      = typename Ext::size_type;
    using rank_type      = typename Ext::rank_type;
    using container_type = Storage;
    using view_type       = ::Kokkos::mdspan<Elem, Ext, Layout>;
    using const_view_type = ::Kokkos::mdspan<Elem const, Ext, Layout>;

    mdarray() = default;

    explicit mdarray(Ext const& ext) : _mapping{ext}, _storage(_mapping.required_span_size()) {}

    template<typename... Ints>
        requires((sizeof...(Ints) > 0) and (std::is_convertible_v<Ints, index_type> and ...))
    explicit mdarray(Ints... dims) : mdarray{Ext{static_cast<index_type>(dims)...}}
    {}

    [[nodiscard]] static constexpr auto rank() noexcept -> rank_type { return Ext::rank(); }
    [[nodiscard]] auto extents() const noexcept -> Ext const& { return _mapping.extents(); }
    [[nodiscard]] auto extent(std::size_t r) const noexcept -> index_type { return extents().extent(r); }
    [[nodiscard]] auto mapping() const noexcept -> mapping_type const& { return _mapping; }
    [[nodiscard]] auto stride(std::size_t r) const -> index_type { return _mapping.stride(r); }

    [[nodiscard]] auto size() const noexcept -> size_type
    {
        auto n = size_type{1};
        for (rank_type r = 0; r < rank(); ++r) { n *= static_cast<size_type>(extent(r)); }
        return n;
    }

    [[nodiscard]] auto data() noexcept -> Elem* { return _storage.data(); }
    [[nodiscard]] auto data() const noexcept -> Elem const* { return _storage.data(); }

    template<typename... Ints>
    [[nodiscard]] auto operator()(Ints... idx) -> Elem&
    {
        return _storage[static_cast<std::size_t>(_mapping(static_cast<index_type>(idx)...))];
    }

    template<typename... Ints>
    [[nodiscard]] auto operator()(Ints... idx) const -> Elem const&
    {
        return _storage[static_cast<std::size_t>(_mapping(static_cast<index_type>(idx)...))];
    }

    [[nodiscard]] auto to_mdspan() -> view_type { return view_type{_storage.data(), _mapping}; }
    [[nodiscard]] auto to_mdspan() const -> const_view_type { return const_view_type{_storage.data(), _mapping}; }

    template<typename E2, typename X2, typename L2, typename A2>
    operator ::Kokkos::mdspan<E2, X2, L2, A2>()
    {
        return ::Kokkos::mdspan<E2, X2, L2, A2>{_storage.data(), _mapping};
    }

    template<typename E2, typename X2, typename L2, typename A2>
    operator ::Kokkos::mdspan<E2, X2, L2, A2>() const
    {
        return ::Kokkos::mdspan<E2, X2, L2, A2>{_storage.data(), _mapping};
    }

private:
    mapping_type _mapping{};
    Storage _storage{};
};

}  // namespace Kokkos::Experimental
