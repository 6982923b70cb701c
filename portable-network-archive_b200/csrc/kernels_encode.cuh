// kernels_encode.cuh -- encode-side kernels (K6/K7 + cipher encrypt + FDAT CRC).  Placeholder until the
// encoders land: the entry points report PNA_E_INTERNAL so nothing silently falls back.
#pragma once
#include "common.cuh"
namespace pna { namespace enc {
struct EncodePlan {};
inline bool init_attributes() { return true; }
inline void destroy(EncodePlan* p) { delete p; }
}}
