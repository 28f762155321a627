// unicode_tables.h -- Unicode character data for the device normalizer (host side).
#pragma once
#include <cstddef>
#include <cstdint>

namespace tmgpu {

struct UnicodeHost {
  const uint16_t* stage1; size_t n_stage1;
  const uint16_t* stage2; size_t n_stage2;
  const uint32_t* dkey; size_t n_decomp;
  const uint16_t* didx;             // n_decomp + 1
  const uint32_t* ddata; size_t n_ddata;
  const uint32_t* lkey; const uint32_t* lval; size_t n_lower;
  const uint32_t* ukey; const uint32_t* uval; size_t n_upper;
  const char* version;
};

const UnicodeHost& unicode_host_tables();

}  // namespace tmgpu
