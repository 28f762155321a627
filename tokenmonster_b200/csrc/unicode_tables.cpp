// unicode_tables.cpp -- host copy of the generated Unicode character data (unicode_tables.inc).
#include "unicode_tables.h"

namespace tmgpu {
namespace {
#include "unicode_tables.inc"
}

const UnicodeHost& unicode_host_tables() {
  static const UnicodeHost t = {
      kUniStage1,    sizeof(kUniStage1) / sizeof(kUniStage1[0]),
      kUniStage2,    sizeof(kUniStage2) / sizeof(kUniStage2[0]),
      kUniDecompKey, sizeof(kUniDecompKey) / sizeof(kUniDecompKey[0]),
      kUniDecompIdx, kUniDecompData, sizeof(kUniDecompData) / sizeof(kUniDecompData[0]),
      kUniLowerKey,  kUniLowerVal,   sizeof(kUniLowerKey) / sizeof(kUniLowerKey[0]),
      kUniUpperKey,  kUniUpperVal,   sizeof(kUniUpperKey) / sizeof(kUniUpperKey[0]),
      TMGPU_UNICODE_VERSION,
  };
  return t;
}

}  // namespace tmgpu
