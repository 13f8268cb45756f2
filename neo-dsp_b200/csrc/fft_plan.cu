// fft_plan.cu -- C ABI of the c2c and r2c/c2r plans (include/neo_b200.h), the drop-in for
// neo::fft::fft_plan (src/neo/fft/reference/c2c_dit2_plan.hpp:22-104) and
// neo::fft::rfft_plan (src/neo/fft/fallback/fallback_rfft_plan.hpp:15-61).
#include "fft_kernels.cuh"
#include "fft_large.cuh"

#include <cmath>
#include <memory>

namespace neo_b200 {

std::atomic<std::uint64_t>& launch_counter()
{
    static std::atomic<std::uint64_t> counter{0};
    return counter;
}

int require_device()
{
    int count             = 0;
    cudaError_t const err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(NEO_B200_ERR_CUDA, "no CUDA device available (%s); this backend has no CPU fallback",
                    err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
    }
    return NEO_B200_OK;
}

namespace {

constexpr size_t k_max_order = 27;  // c2c_dit2_plan::max_order(), c2c_dit2_plan.hpp:59-62

template<typename T>
struct c2c_engine
{
    int order;
    fft_tables<T> tables;        // single-CTA path
    large_fft<T> large;          // four-step path above max_cta_logm
    bool use_large{false};

    int init(int order_, cudaStream_t stream)
    {
        order     = order_;
        use_large = order > max_cta_logm<T>();
        if (use_large) { return large.init(order, stream); }
        return tables.build(order, false, stream);
    }

    int exec(cx<T> const* in, cx<T>* out, size_t batch, int direction, cudaStream_t stream)
    {
        if (use_large) { return large.exec(in, out, batch, direction, stream); }
        int status = NEO_B200_ERR_UNSUPPORTED;
        NEO_DISPATCH_LOGM(T, order, {
            if constexpr (LOGM <= max_cta_logm<T>()) {
                status = direction < 0 ? launch_c2c<T, LOGM, -1>(in, out, tables.tw(), batch, stream)
                                       : launch_c2c<T, LOGM, +1>(in, out, tables.tw(), batch, stream);
            }
        });
        if (status == NEO_B200_ERR_UNSUPPORTED) { return fail(status, "c2c order %d not supported", order); }
        return status;
    }
};

template<typename T>
struct rfft_engine
{
    int order;                   // real size N = 2^order, complex half size M = N/2
    fft_tables<T> tables;
    large_rfft<T> large;
    bool use_large{false};

    int init(int order_, cudaStream_t stream)
    {
        order     = order_;
        use_large = order - 1 > max_cta_logm<T>();
        if (order == 0) { return NEO_B200_OK; }
        if (use_large) { return large.init(order, stream); }
        return tables.build(order - 1, true, stream);
    }

    int forward(T const* in, cx<T>* out, size_t batch, cudaStream_t stream)
    {
        if (order == 0) { return large_rfft<T>::size_one_forward(in, out, batch, stream); }
        if (use_large) { return large.forward(in, out, batch, stream); }
        int status = NEO_B200_ERR_UNSUPPORTED;
        NEO_DISPATCH_LOGM(T, order - 1, {
            if constexpr (LOGM <= max_cta_logm<T>()) {
                status = launch_r2c<T, LOGM>(r2c_plain_io<T, LOGM>{in, out}, tables.tw(), tables.rtw(), batch, stream);
            }
        });
        if (status == NEO_B200_ERR_UNSUPPORTED) { return fail(status, "rfft order %d not supported", order); }
        return status;
    }

    int backward(cx<T> const* in, size_t row_len, T* out, size_t batch, cudaStream_t stream)
    {
        if (order == 0) { return large_rfft<T>::size_one_backward(in, row_len, out, batch, stream); }
        if (use_large) { return large.backward(in, row_len, out, batch, stream); }
        int status = NEO_B200_ERR_UNSUPPORTED;
        NEO_DISPATCH_LOGM(T, order - 1, {
            if constexpr (LOGM <= max_cta_logm<T>()) {
                status = launch_c2r<T, LOGM>(c2r_plain_io<T, LOGM>{in, out, row_len}, tables.tw(), tables.rtw(), batch, stream);
            }
        });
        if (status == NEO_B200_ERR_UNSUPPORTED) { return fail(status, "irfft order %d not supported", order); }
        return status;
    }
};

}  // namespace
}  // namespace neo_b200

using namespace neo_b200;

struct neo_b200_fft_plan
{
    size_t order;
    int dtype;
    int device;
    stream_ref stream;
    c2c_engine<float> f32;
    c2c_engine<double> f64;
    device_buffer staging_in, staging_out;
};

struct neo_b200_rfft_plan
{
    size_t order;
    int dtype;
    int device;
    stream_ref stream;
    rfft_engine<float> f32;
    rfft_engine<double> f64;
    device_buffer staging_in, staging_out;
};

extern "C" {

const char* neo_b200_last_error(void) { return last_error_slot().c_str(); }

const char* neo_b200_version(void) { return "neo_b200 0.1 (sm_100a)"; }

int neo_b200_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return count;
}

int neo_b200_set_device(int device)
{
    NEO_CUDA_TRY(cudaSetDevice(device));
    return NEO_B200_OK;
}

int neo_b200_kernel_launches(uint64_t* count)
{
    if (count == nullptr) { return fail(NEO_B200_ERR_INVALID, "count is null"); }
    *count = launch_counter().load();
    return NEO_B200_OK;
}

size_t neo_b200_fft_max_order(void) { return k_max_order; }

size_t neo_b200_next_order(size_t size)
{
    size_t order = 0;
    while ((size_t(1) << order) < size) { ++order; }
    return order;
}

// ---- c2c ---------------------------------------------------------------------------------------------------------------
int neo_b200_fft_plan_create(neo_b200_fft_plan** plan, size_t order, int dtype)
{
    if (plan == nullptr) { return fail(NEO_B200_ERR_INVALID, "plan is null"); }
    *plan = nullptr;
    if (dtype != NEO_B200_F32 && dtype != NEO_B200_F64) { return fail(NEO_B200_ERR_INVALID, "bad dtype %d", dtype); }
    // same contract as c2c_dit2_plan::check_order (c2c_dit2_plan.hpp:98-104)
    if (order > k_max_order) { return fail(NEO_B200_ERR_UNSUPPORTED, "neo_b200: unsupported order '%zu'", order); }
    NEO_TRY(require_device());
    auto p    = std::unique_ptr<neo_b200_fft_plan>(new (std::nothrow) neo_b200_fft_plan{});
    if (!p) { return fail(NEO_B200_ERR_ALLOC, "out of host memory"); }
    p->order = order;
    p->dtype = dtype;
    NEO_CUDA_TRY(cudaGetDevice(&p->device));
    NEO_TRY(p->stream.create());
    if (dtype == NEO_B200_F32) { NEO_TRY(p->f32.init(static_cast<int>(order), p->stream.stream)); }
    else { NEO_TRY(p->f64.init(static_cast<int>(order), p->stream.stream)); }
    *plan = p.release();
    return NEO_B200_OK;
}

void neo_b200_fft_plan_destroy(neo_b200_fft_plan* plan)
{
    if (plan == nullptr) { return; }
    cudaStreamSynchronize(plan->stream.stream);
    delete plan;
}

size_t neo_b200_fft_plan_order(neo_b200_fft_plan const* plan) { return plan != nullptr ? plan->order : 0; }
size_t neo_b200_fft_plan_size(neo_b200_fft_plan const* plan) { return plan != nullptr ? size_t(1) << plan->order : 0; }

int neo_b200_fft_plan_set_stream(neo_b200_fft_plan* plan, void* cuda_stream)
{
    if (plan == nullptr) { return fail(NEO_B200_ERR_INVALID, "plan is null"); }
    plan->stream.adopt(cuda_stream);
    return NEO_B200_OK;
}

int neo_b200_fft_plan_synchronize(neo_b200_fft_plan* plan)
{
    if (plan == nullptr) { return fail(NEO_B200_ERR_INVALID, "plan is null"); }
    NEO_CUDA_TRY(cudaStreamSynchronize(plan->stream.stream));
    return NEO_B200_OK;
}

static int fft_exec_device(neo_b200_fft_plan* plan, void const* in, void* out, size_t batch, int direction)
{
    if (plan->dtype == NEO_B200_F32) {
        return plan->f32.exec(static_cast<float2 const*>(in), static_cast<float2*>(out), batch, direction, plan->stream.stream);
    }
    return plan->f64.exec(static_cast<double2 const*>(in), static_cast<double2*>(out), batch, direction, plan->stream.stream);
}

int neo_b200_fft_exec(neo_b200_fft_plan* plan, void const* in, void* out, size_t batch, int direction, int memspace)
{
    if (plan == nullptr || in == nullptr || out == nullptr) { return fail(NEO_B200_ERR_INVALID, "null argument"); }
    if (direction != NEO_B200_FORWARD && direction != NEO_B200_BACKWARD) {
        return fail(NEO_B200_ERR_INVALID, "direction must be -1 (forward) or +1 (backward)");
    }
    if (batch == 0) { return NEO_B200_OK; }
    NEO_CUDA_TRY(cudaSetDevice(plan->device));
    if (memspace == NEO_B200_DEVICE) { return fft_exec_device(plan, in, out, batch, direction); }

    // HOST: stage through device memory in chunks, synchronous semantics like the reference
    size_t const row   = (size_t(1) << plan->order) * 2 * elem_size(plan->dtype);
    size_t const chunk = std::max<size_t>(1, std::min(batch, (size_t(256) << 20) / row));
    NEO_TRY(plan->staging_in.reserve(chunk * row));
    cudaStream_t const s = plan->stream.stream;
    for (size_t first = 0; first < batch; first += chunk) {
        size_t const n = std::min(chunk, batch - first);
        NEO_CUDA_TRY(cudaMemcpyAsync(plan->staging_in.ptr, static_cast<char const*>(in) + first * row, n * row, cudaMemcpyHostToDevice, s));
        NEO_TRY(fft_exec_device(plan, plan->staging_in.ptr, plan->staging_in.ptr, n, direction));
        NEO_CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(out) + first * row, plan->staging_in.ptr, n * row, cudaMemcpyDeviceToHost, s));
    }
    NEO_CUDA_TRY(cudaStreamSynchronize(s));
    return NEO_B200_OK;
}

int neo_b200_fft_exec_strided(
    neo_b200_fft_plan* plan, void const* in, ptrdiff_t in_stride, void* out, ptrdiff_t out_stride, int direction)
{
    if (plan == nullptr || in == nullptr || out == nullptr) { return fail(NEO_B200_ERR_INVALID, "null argument"); }
    NEO_CUDA_TRY(cudaSetDevice(plan->device));
    size_t const n    = size_t(1) << plan->order;
    size_t const elem = 2 * elem_size(plan->dtype);
    NEO_TRY(plan->staging_in.reserve(n * elem));
    cudaStream_t const s = plan->stream.stream;
    // strided gather / scatter with 2-D copies: n rows of one complex element, pitch = stride
    if (in_stride < 1 || out_stride < 1) { return fail(NEO_B200_ERR_INVALID, "strides must be positive"); }
    NEO_CUDA_TRY(cudaMemcpy2DAsync(plan->staging_in.ptr, elem, in, size_t(in_stride) * elem, elem, n, cudaMemcpyHostToDevice, s));
    NEO_TRY(fft_exec_device(plan, plan->staging_in.ptr, plan->staging_in.ptr, 1, direction));
    NEO_CUDA_TRY(cudaMemcpy2DAsync(out, size_t(out_stride) * elem, plan->staging_in.ptr, elem, elem, n, cudaMemcpyDeviceToHost, s));
    NEO_CUDA_TRY(cudaStreamSynchronize(s));
    return NEO_B200_OK;
}

// ---- rfft ----------------------------------------------------------------------------------------------------------------
int neo_b200_rfft_plan_create(neo_b200_rfft_plan** plan, size_t order, int dtype)
{
    if (plan == nullptr) { return fail(NEO_B200_ERR_INVALID, "plan is null"); }
    *plan = nullptr;
    if (dtype != NEO_B200_F32 && dtype != NEO_B200_F64) { return fail(NEO_B200_ERR_INVALID, "bad dtype %d", dtype); }
    // fallback_rfft_plan owns an fft_plan of the same order (fallback_rfft_plan.hpp:58), so the same limit applies
    if (order > k_max_order) { return fail(NEO_B200_ERR_UNSUPPORTED, "neo_b200: unsupported order '%zu'", order); }
    NEO_TRY(require_device());
    auto p = std::unique_ptr<neo_b200_rfft_plan>(new (std::nothrow) neo_b200_rfft_plan{});
    if (!p) { return fail(NEO_B200_ERR_ALLOC, "out of host memory"); }
    p->order = order;
    p->dtype = dtype;
    NEO_CUDA_TRY(cudaGetDevice(&p->device));
    NEO_TRY(p->stream.create());
    if (dtype == NEO_B200_F32) { NEO_TRY(p->f32.init(static_cast<int>(order), p->stream.stream)); }
    else { NEO_TRY(p->f64.init(static_cast<int>(order), p->stream.stream)); }
    *plan = p.release();
    return NEO_B200_OK;
}

void neo_b200_rfft_plan_destroy(neo_b200_rfft_plan* plan)
{
    if (plan == nullptr) { return; }
    cudaStreamSynchronize(plan->stream.stream);
    delete plan;
}

size_t neo_b200_rfft_plan_order(neo_b200_rfft_plan const* plan) { return plan != nullptr ? plan->order : 0; }
size_t neo_b200_rfft_plan_size(neo_b200_rfft_plan const* plan) { return plan != nullptr ? size_t(1) << plan->order : 0; }

int neo_b200_rfft_plan_set_stream(neo_b200_rfft_plan* plan, void* cuda_stream)
{
    if (plan == nullptr) { return fail(NEO_B200_ERR_INVALID, "plan is null"); }
    plan->stream.adopt(cuda_stream);
    return NEO_B200_OK;
}

int neo_b200_rfft_plan_synchronize(neo_b200_rfft_plan* plan)
{
    if (plan == nullptr) { return fail(NEO_B200_ERR_INVALID, "plan is null"); }
    NEO_CUDA_TRY(cudaStreamSynchronize(plan->stream.stream));
    return NEO_B200_OK;
}

static int rfft_device(neo_b200_rfft_plan* plan, void const* in, void* out, size_t batch)
{
    cudaStream_t const s = plan->stream.stream;
    if (plan->dtype == NEO_B200_F32) { return plan->f32.forward(static_cast<float const*>(in), static_cast<float2*>(out), batch, s); }
    return plan->f64.forward(static_cast<double const*>(in), static_cast<double2*>(out), batch, s);
}

static int irfft_device(neo_b200_rfft_plan* plan, void const* in, size_t row_len, void* out, size_t batch)
{
    cudaStream_t const s = plan->stream.stream;
    if (plan->dtype == NEO_B200_F32) {
        return plan->f32.backward(static_cast<float2 const*>(in), row_len, static_cast<float*>(out), batch, s);
    }
    return plan->f64.backward(static_cast<double2 const*>(in), row_len, static_cast<double*>(out), batch, s);
}

int neo_b200_rfft_exec(neo_b200_rfft_plan* plan, void const* in, void* out, size_t batch, int memspace)
{
    if (plan == nullptr || in == nullptr || out == nullptr) { return fail(NEO_B200_ERR_INVALID, "null argument"); }
    if (batch == 0) { return NEO_B200_OK; }
    NEO_CUDA_TRY(cudaSetDevice(plan->device));
    if (memspace == NEO_B200_DEVICE) { return rfft_device(plan, in, out, batch); }

    size_t const n       = size_t(1) << plan->order;
    size_t const in_row  = n * elem_size(plan->dtype);
    size_t const out_row = (n / 2 + 1) * 2 * elem_size(plan->dtype);
    size_t const chunk   = std::max<size_t>(1, std::min(batch, (size_t(256) << 20) / in_row));
    NEO_TRY(plan->staging_in.reserve(chunk * in_row));
    NEO_TRY(plan->staging_out.reserve(chunk * out_row));
    cudaStream_t const s = plan->stream.stream;
    for (size_t first = 0; first < batch; first += chunk) {
        size_t const cnt = std::min(chunk, batch - first);
        NEO_CUDA_TRY(cudaMemcpyAsync(plan->staging_in.ptr, static_cast<char const*>(in) + first * in_row, cnt * in_row, cudaMemcpyHostToDevice, s));
        NEO_TRY(rfft_device(plan, plan->staging_in.ptr, plan->staging_out.ptr, cnt));
        NEO_CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(out) + first * out_row, plan->staging_out.ptr, cnt * out_row, cudaMemcpyDeviceToHost, s));
    }
    NEO_CUDA_TRY(cudaStreamSynchronize(s));
    return NEO_B200_OK;
}

int neo_b200_irfft_exec(neo_b200_rfft_plan* plan, void const* in, size_t in_row_len, void* out, size_t batch, int memspace)
{
    if (plan == nullptr || in == nullptr || out == nullptr) { return fail(NEO_B200_ERR_INVALID, "null argument"); }
    size_t const n = size_t(1) << plan->order;
    if (in_row_len < n / 2 + 1) { return fail(NEO_B200_ERR_INVALID, "irfft needs at least N/2+1 = %zu bins per row, got %zu", n / 2 + 1, in_row_len); }
    if (batch == 0) { return NEO_B200_OK; }
    NEO_CUDA_TRY(cudaSetDevice(plan->device));
    if (memspace == NEO_B200_DEVICE) { return irfft_device(plan, in, in_row_len, out, batch); }

    size_t const in_row  = in_row_len * 2 * elem_size(plan->dtype);
    size_t const out_row = n * elem_size(plan->dtype);
    size_t const chunk   = std::max<size_t>(1, std::min(batch, (size_t(256) << 20) / in_row));
    NEO_TRY(plan->staging_in.reserve(chunk * in_row));
    NEO_TRY(plan->staging_out.reserve(chunk * out_row));
    cudaStream_t const s = plan->stream.stream;
    for (size_t first = 0; first < batch; first += chunk) {
        size_t const cnt = std::min(chunk, batch - first);
        NEO_CUDA_TRY(cudaMemcpyAsync(plan->staging_in.ptr, static_cast<char const*>(in) + first * in_row, cnt * in_row, cudaMemcpyHostToDevice, s));
        NEO_TRY(irfft_device(plan, plan->staging_in.ptr, in_row_len, plan->staging_out.ptr, cnt));
        NEO_CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(out) + first * out_row, plan->staging_out.ptr, cnt * out_row, cudaMemcpyDeviceToHost, s));
    }
    NEO_CUDA_TRY(cudaStreamSynchronize(s));
    return NEO_B200_OK;
}

}  // extern "C"
