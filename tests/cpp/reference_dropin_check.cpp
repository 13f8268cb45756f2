// reference_dropin_check.cpp -- COMPILE-ONLY proof that the facade satisfies the reference's duck types: neo's own free functions
// (fft/fft.hpp:56-90, fft/rfft.hpp:28-39) and its mdspan vocabulary are instantiated with the neo::b200 classes.
// Built only where /root/reference exists (tests/cpp/Makefile), against the unmodified headers + oracle/shim.
#include <neo/complex.hpp>
#include <neo/container/mdspan.hpp>
#include <neo/fft.hpp>

#include "../../include/neo_b200.hpp"

template<typename Complex>
auto instantiate_c2c() -> void
{
    auto plan = neo::b200::fft_plan<Complex>{neo::fft::from_order, 4};
    auto buf  = stdex::mdarray<Complex, stdex::dextents<std::size_t, 1>>{plan.size()};
    auto out  = stdex::mdarray<Complex, stdex::dextents<std::size_t, 1>>{plan.size()};
    neo::fft::fft(plan, buf.to_mdspan());                   // in place
    neo::fft::ifft(plan, buf.to_mdspan());
    neo::fft::fft(plan, buf.to_mdspan(), out.to_mdspan());  // picks the 3-argument overload through `requires`
    neo::fft::ifft(plan, buf.to_mdspan(), out.to_mdspan());
    static_assert(neo::b200::fft_plan<Complex>::max_order() == neo::fft::fft_plan<Complex>::max_order());
}

template<typename Float>
auto instantiate_r2c() -> void
{
    using Complex = std::complex<Float>;
    auto plan     = neo::b200::rfft_plan<Float, Complex>{neo::fft::from_order, 4};
    auto real     = stdex::mdarray<Float, stdex::dextents<std::size_t, 1>>{plan.size()};
    auto cplx     = stdex::mdarray<Complex, stdex::dextents<std::size_t, 1>>{plan.size()};
    neo::fft::rfft(plan, real.to_mdspan(), cplx.to_mdspan());
    neo::fft::irfft(plan, cplx.to_mdspan(), real.to_mdspan());
}

template<typename Float>
auto instantiate_split() -> void
{
    auto plan = neo::b200::split_fft_plan<Float>{neo::fft::from_order, 4};
    auto re   = stdex::mdarray<Float, stdex::dextents<std::size_t, 1>>{plan.size()};
    auto im   = stdex::mdarray<Float, stdex::dextents<std::size_t, 1>>{plan.size()};
    auto x    = neo::split_complex{re.to_mdspan(), im.to_mdspan()};
    neo::fft::fft(plan, x);        // fft/split_fft.hpp:38-43
    neo::fft::ifft(plan, x);
    neo::fft::fft(plan, x, x);     // out-of-place overload, fft/split_fft.hpp:46-51
}

template<typename Float>
auto instantiate_fft_convolver() -> void
{
    auto conv   = neo::b200::fft_convolver<Float>{100, 31};  // convolution/fft_convolver.hpp:21
    auto signal = stdex::mdarray<Float, stdex::dextents<std::size_t, 1>>{100};
    auto patch  = stdex::mdarray<Float, stdex::dextents<std::size_t, 1>>{31};
    auto output = stdex::mdarray<Float, stdex::dextents<std::size_t, 1>>{conv.output_size()};
    conv(signal.to_mdspan(), patch.to_mdspan(), output.to_mdspan());  // :39
}

template<typename Float>
auto instantiate_dct2() -> void
{
    auto plan = neo::b200::dct2_plan<Float>{neo::fft::from_order, 3};  // fft/dct_test.cpp:17-22
    auto buf  = stdex::mdarray<Float, stdex::dextents<std::size_t, 1>>{8};
    plan(buf.to_mdspan());
    (void)plan.order();
    (void)plan.size();
}

template<typename Complex>
auto instantiate_dft() -> void
{
    auto plan = neo::b200::dft_plan<Complex>{21};  // fft/dft_test.cpp:24-29: Plan{size}
    auto buf  = stdex::mdarray<Complex, stdex::dextents<std::size_t, 1>>{21};
    neo::fft::dft(plan, buf.to_mdspan());   // fft/dft.hpp: dft(plan, x)
    neo::fft::idft(plan, buf.to_mdspan());
    (void)plan.size();
}

template<typename Complex>
auto instantiate_convolver() -> void
{
    using Float = typename Complex::value_type;
    auto conv   = neo::b200::upols_convolver<Complex>{};
    auto h      = stdex::mdarray<Complex, stdex::dextents<std::size_t, 2>>{3, 129};
    auto block  = stdex::mdarray<Float, stdex::dextents<std::size_t, 1>>{128};
    conv.filter(h.to_mdspan());  // uniform_partitioned_convolver.hpp:24
    conv(block.to_mdspan());     // uniform_partitioned_convolver.hpp:25
    auto ola = neo::b200::upola_convolver<Complex>{};
    ola.filter(h.to_mdspan());
    ola(block.to_mdspan());
    auto ola2 = neo::b200::upola_convolver_v2<Complex>{};  // overlap_add_convolver.hpp:32-33
    ola2.filter(h.to_mdspan());
    ola2(block.to_mdspan());
    auto sparse = neo::b200::sparse_upols_convolver<Complex>{};  // sparse_convolver.hpp:14-22, test :57-61
    sparse.filter(h.to_mdspan(), [](auto, auto, auto) { return true; });
    sparse(block.to_mdspan());
    auto sparse_ola = neo::b200::sparse_upola_convolver<Complex>{};
    sparse_ola.filter(h.to_mdspan(), [](auto, auto, auto) { return true; });
    sparse_ola(block.to_mdspan());
    auto split = neo::b200::split_upols_convolver<Complex>{};  // dense_convolver.hpp:32-41
    split.filter(h.to_mdspan());
    split(block.to_mdspan());
    auto split_ola = neo::b200::split_upola_convolver<Complex>{};
    split_ola.filter(h.to_mdspan());
    split_ola(block.to_mdspan());
}

auto instantiate_all() -> void
{
    instantiate_c2c<std::complex<float>>();
    instantiate_c2c<std::complex<double>>();
    instantiate_c2c<neo::scalar_complex<float>>();
    instantiate_split<float>();
    instantiate_split<double>();
    instantiate_r2c<float>();
    instantiate_r2c<double>();
    instantiate_fft_convolver<float>();
    instantiate_fft_convolver<double>();
    instantiate_dct2<float>();
    instantiate_dct2<double>();
    instantiate_dft<std::complex<float>>();
    instantiate_dft<std::complex<double>>();
    instantiate_convolver<std::complex<float>>();
    instantiate_convolver<std::complex<double>>();
}
