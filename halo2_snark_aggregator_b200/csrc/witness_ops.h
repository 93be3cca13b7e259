// Op records exchanged between the host-side recorder (witness_recorder.cu) and the row-expansion
// kernel (witness.cu).  One record = one fixed row recipe of halo2-ecc-circuit-lib
// (SURVEY.md 8a rows W1-W5); `row` is the first advice row it owns, so records are independent
// and the expansion is embarrassingly parallel.  The device groups the records by opcode (index array), so every
// recipe runs as its own kernel (its own register budget, no divergence).
#pragma once
#include <cstdint>

namespace h2agg {

enum WitnessOpcode : uint32_t {
  WOP_RAW128 = 0,   // aux = nrows (1..3); v[10*r + 2*c .. +2) = cell (r, c) as a 128-bit integer
  WOP_RAW256 = 1,   // one row; v[4*c .. +4) = cell c as a canonical 256-bit integer
  WOP_NATIVE = 2,   // a = v[0..8): [native(a), a0, a1, a2, a3]                       five/integer_chip.rs:595-621
  WOP_REDUCE = 3,   // a = v[0..8), rem = v[8..16); flags bit0 = native(a) cached     :483-581
  WOP_ISZERO = 4,   // a = v[0..8) (already reduced); flags bit0 = native cached      :53-102, 796-806
  WOP_MULEQ = 5,    // x = v[0..8), y = v[8..16), z = v[16..24): x*y = d*p + z         :104-320, 709-782
                    // flags: bit0/1/2 native cached for x/y/z, bit3 square (y is x),
                    //        bit4 the freshly assign_w'd integer is y (div) instead of z (mul)
  WOP_COUNT = 6,
};

struct alignas(16) WitnessOp {
  uint32_t opcode;
  uint32_t row;
  uint32_t flags;
  uint32_t aux;
  uint64_t v[30];
};
static_assert(sizeof(WitnessOp) == 256, "record is 256 bytes");

}  // namespace h2agg
