// phx_sc_wire.h -- compact wire format of one supply-chain env-step (host-buffer path).
//
// The lean output row of an env-step -- obs [stock / max_stock, sales / cap, missed / cap] as
// float32, reward float32(sales - 0.1 stock), all_done {0, truncated} -- is a function of three
// small integers and two flags.  Across PCIe it travels as ONE 32-bit word (4 B instead of 18 B)
//     bits  0..15  stock + 2^15, the stock the reward saw (before an auto-reset)
//     bits 16..22  sales          (0..127)
//     bits 23..29  missed_sales   (0..127)
//     bit  30      truncations["__all__"] (the episode's last step)
//     bit  31      the env was auto-reset in this step: the observation shows stock = 0
// and is expanded on the host (phx_sc_wire.cpp) with the same correctly rounded quotients the
// kernel computes (fam_supply_chain.cu sc_ratio; the reference's float64 division + float32
// cast, supply_chain.py:124-134,144-147).  Values outside the fields (possible only with
// out-of-distribution actions) raise the launch's overflow flag and the float planes are used.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SCW_HD __host__ __device__ __forceinline__
#else
#define SCW_HD inline
#endif

namespace phx {

constexpr int SCW_STOCK_BIAS = 32768;
constexpr uint32_t SCW_FIELD_MAX = 127;

SCW_HD uint32_t scw_pack(int stock_pre, int sales, int missed, bool truncated, bool was_reset) {
  return ((uint32_t)(stock_pre + SCW_STOCK_BIAS) & 0xFFFFu) | (((uint32_t)sales & 127u) << 16) |
         (((uint32_t)missed & 127u) << 23) | ((uint32_t)truncated << 30) |
         ((uint32_t)was_reset << 31);
}

struct ScWireParams {
  int32_t max_stock;  // obs[0] denominator (SHOP_MAX_STOCK)
  int32_t cap;        // obs[1], obs[2] denominator (n_customers * CUSTOMER_MAX_ORDER_SIZE)
};

class HostPool;
// Expands `n` wire words into obs float[n,3], reward float[n], all_done uint8[n,2].
void sc_wire_expand(HostPool& pool, const ScWireParams& p, const uint32_t* wire, size_t n,
                    float* obs, float* reward, uint8_t* all_done);
// the scalar reference form of one word (also the tail / non-AVX2 path)
void sc_wire_expand_scalar(const ScWireParams& p, const uint32_t* wire, size_t n, float* obs,
                           float* reward, uint8_t* all_done);

}  // namespace phx
