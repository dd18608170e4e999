// TEST INFRASTRUCTURE: compiles the device headers' host (PTX-emulation) back-end so the
// algorithms in zolt_b200/csrc/*.cuh can be validated against the oracle without a GPU.
// Never linked into the product library.
#include "../../zolt_b200/csrc/field.cuh"
#include <cstddef>
#include <cstring>
using namespace zolt;

template <class P>
static void binop(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    Fp<P> x, y, r;
    memcpy(x.v, a + 8 * i, 32);
    memcpy(y.v, b + 8 * i, 32);
    switch (op) {
      case 0: r = fp_mul<P>(x, y); break;
      case 1: r = fp_add<P>(x, y); break;
      case 2: r = fp_sub<P>(x, y); break;
      case 3: r = fp_inv<P>(x); break;
      case 4: r = fp_from_mont<P>(x); break;
      case 5: r = fp_to_mont<P>(x); break;
      case 6: r = fp_neg<P>(x); break;
      default: r = Fp<P>::zero();
    }
    memcpy(out + 8 * i, r.v, 32);
  }
}

extern "C" void shim_field_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  if (field == 0) binop<FrParams>(op, a, b, out, n);
  else binop<FqParams>(op, a, b, out, n);
}
