// ads/lin/tensor.hpp -- ads::lin::tensor<T, Rank>: owning column-major (first index fastest) N-d
// array with the interface of the reference (include/ads/lin/tensor/tensor.hpp:14-50, base.hpp,
// ordering include/ads/util/multi_array/ordering/reverse.hpp:28-31), plus a lazily synchronised
// device mirror: the hot path works on the device copy; element access from host code downloads it
// first (and invalidates the device copy when the access can write).
#ifndef ADSB_ADS_LIN_TENSOR_HPP
#define ADSB_ADS_LIN_TENSOR_HPP

#include <algorithm>
#include <array>
#include <cstddef>
#include <vector>

#include "ads/device.hpp"

namespace ads::lin {

template <typename T, std::size_t Rank>
class tensor {
public:
    using size_array = std::array<int, Rank>;

    explicit tensor(const size_array& sizes) : sizes_{sizes}, data_(count(sizes)) { }

    // Value semantics of the reference tensor (a plain std::vector owner): a copy holds the source's CURRENT
    // values -- downloaded first when the device copy is the newer one -- and starts without a device mirror
    // (the mirror's managed buffer belongs to the source alone).  Moves carry the mirror along, so
    // std::swap(u, u_prev) keeps both tensors resident.
    tensor(const tensor& o) : sizes_{o.sizes_}, data_(o.host_data()) { }
    tensor& operator=(const tensor& o) {
        if (this != &o) {
            const bool same_shape = sizes_ == o.sizes_;
            data_ = o.host_data();
            sizes_ = o.sizes_;
            if (!same_shape) release();  // the managed buffer was sized for the old shape
            m_.host_valid = true;        // an attached destination stays attached; its device copy is stale now
            m_.dev_valid = false;
        }
        return *this;
    }
    tensor(tensor&& o) noexcept : sizes_{o.sizes_}, data_(std::move(o.data_)), m_(std::move(o.m_)) {
        o.m_ = device::mirror{};
    }
    tensor& operator=(tensor&& o) noexcept {
        if (this != &o) {
            release();
            sizes_ = o.sizes_;
            data_ = std::move(o.data_);
            m_ = std::move(o.m_);
            o.m_ = device::mirror{};
        }
        return *this;
    }
    ~tensor() { release(); }

    int size() const { return static_cast<int>(data_.size()); }
    int size(int dim) const { return sizes_[dim]; }
    const size_array& sizes() const { return sizes_; }

    template <typename... Idx>
    T& operator()(Idx... idx) {
        host_for_write();
        return data_[linear(idx...)];
    }
    template <typename... Idx>
    const T& operator()(Idx... idx) const {
        host_for_read();
        return data_[linear(idx...)];
    }

    T* data() {
        host_for_write();
        return data_.data();
    }
    const T* data() const {
        host_for_read();
        return data_.data();
    }

    void fill_with_zeros() {
        std::fill(data_.begin(), data_.end(), T{});
        m_.host_valid = true;
        m_.dev_valid = false;
    }

    // ---- device mirror (used by ads::simulation_Nd / ads::ads_solve)
    bool attached() const { return m_.buf >= 0; }
    void attach(std::shared_ptr<device::context> ctx) {
        if (attached() && m_.ctx == ctx) return;
        host_for_read();
        m_.ctx = std::move(ctx);
        m_.buf = m_.ctx->new_buffer();
        m_.dev_valid = false;
    }
    const std::shared_ptr<device::context>& context() const { return m_.ctx; }
    int device_buffer() const { return m_.buf; }
    void to_device() const {
        if (!m_.dev_valid) {
            device::check(adsb_upload(m_.ctx->handle(), m_.buf, data_.data()));
            m_.dev_valid = true;
        }
    }
    void device_written() {
        m_.dev_valid = true;
        m_.host_valid = false;
    }

private:
    const std::vector<T>& host_data() const {
        host_for_read();
        return data_;
    }
    void release() noexcept {  // hand the managed buffer id back to the context
        if (m_.buf >= 0 && m_.ctx) m_.ctx->free_buffer(m_.buf);
        m_ = device::mirror{};
    }
    static std::size_t count(const size_array& s) {
        std::size_t n = 1;
        for (int v : s) n *= static_cast<std::size_t>(v);
        return n;
    }
    template <typename... Idx>
    std::size_t linear(Idx... idx) const {
        static_assert(sizeof...(Idx) == Rank, "wrong number of indices");
        const int i[] = {idx...};
        std::size_t lin = 0;
        for (std::size_t d = Rank; d-- > 0;) lin = lin * sizes_[d] + i[d];
        return lin;
    }
    void host_for_read() const {
        if (!m_.host_valid) {
            device::check(adsb_download(m_.ctx->handle(), m_.buf, const_cast<T*>(data_.data())));
            m_.host_valid = true;
        }
    }
    void host_for_write() {
        host_for_read();
        m_.dev_valid = false;
    }

    size_array sizes_;
    std::vector<T> data_;
    mutable device::mirror m_;
};

template <typename T, std::size_t Rank>
void zero(tensor<T, Rank>& t) {  // include/ads/lin/tensor/tensor.hpp:44-50
    t.fill_with_zeros();
}

// Non-owning view with the same indexing (include/ads/lin/tensor/view.hpp, as_tensor): host memory only.
template <typename T, std::size_t Rank>
class tensor_view {
public:
    using size_array = std::array<int, Rank>;
    tensor_view(T* data, const size_array& sizes) : data_{data}, sizes_{sizes} { }
    T* data() const { return data_; }
    int size(int dim) const { return sizes_[dim]; }
    const size_array& sizes() const { return sizes_; }
    int size() const {
        int n = 1;
        for (int s : sizes_) n *= s;
        return n;
    }
    template <typename... Idx>
    T& operator()(Idx... idx) const {
        static_assert(sizeof...(Idx) == Rank, "wrong number of indices");
        const int ix[Rank] = {static_cast<int>(idx)...};
        std::size_t lin = 0;
        for (std::size_t d = Rank; d-- > 0;) lin = lin * sizes_[d] + ix[d];  // first index fastest
        return data_[lin];
    }

private:
    T* data_;
    size_array sizes_;
};
template <typename T, std::size_t Rank>
tensor_view<T, Rank> as_tensor(T* data, const std::array<int, Rank>& sizes) {
    return {data, sizes};
}

using vector = tensor<double, 1>;

}  // namespace ads::lin

#endif
