// unicode_dev.h -- device-side view of the Unicode character data and the normalizer's arguments
// (shared by the host code in abi*.cu and the kernels in unicode_normalize.cuh).
#pragma once
#include <cstdint>

namespace tmgpu {

struct UnicodeDev {
  const uint16_t* stage1;   // cp >> 7 -> block
  const uint16_t* stage2;   // block * 128 + (cp & 127) -> property word
  const uint32_t* dkey;     // sorted code points with a canonical decomposition (Hangul excluded)
  const uint16_t* didx;     // n_decomp + 1 offsets into ddata
  const uint32_t* ddata;    // full NFD expansions
  const uint32_t* lkey;     // simple lower-case mapping, sorted keys
  const uint32_t* lval;
  const uint32_t* ukey;     // simple upper-case mapping
  const uint32_t* uval;
  uint32_t n_decomp, n_lower, n_upper;
};

struct NormalizeArgs {
  UnicodeDev U;
  const uint8_t* data;        // raw strings
  const int64_t* off;         // raw begin offsets (n + 1, or n with `ends`)
  const int64_t* ends;        // optional raw end offsets
  const uint32_t* list;       // strings to normalize
  const uint32_t* n_list;     // how many (device scalar)
  uint8_t* scratch;           // normalized text
  const int64_t* n_begin;     // per list item: where its text starts in scratch (capacity 4*len+16)
  int64_t* n_end;             // per list item: end of its text, or n_begin - 1 if refused
  uint32_t normflag, capcode;
};

}  // namespace tmgpu
