// Library-level entry points of the cirq_b200 C-ABI (include/cirq_b200.h).
#include "b2q_common.cuh"

#include <mutex>

namespace b2q {

std::atomic<uint64_t> g_launch_count{0};

char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

namespace {
struct DeviceScratch {
  void* ptr = nullptr;
  size_t bytes = 0;
};
std::mutex g_scratch_mutex;
DeviceScratch g_scratch[64];
}  // namespace

void* workspace(size_t bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    set_error(B2Q_ERR_CUDA, "cudaGetDevice failed");
    return nullptr;
  }
  std::lock_guard<std::mutex> lock(g_scratch_mutex);
  DeviceScratch& s = g_scratch[dev];
  if (s.bytes >= bytes) return s.ptr;
  size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
  if (s.ptr != nullptr) {
    cudaDeviceSynchronize();
    cudaFree(s.ptr);
    s.ptr = nullptr;
    s.bytes = 0;
  }
  if (cudaMalloc(&s.ptr, want) != cudaSuccess) {
    set_error(B2Q_ERR_CUDA, "cudaMalloc of %zu scratch bytes failed", want);
    s.ptr = nullptr;
    return nullptr;
  }
  s.bytes = want;
  return s.ptr;
}

}  // namespace b2q

extern "C" int b2q_version(void) { return 1000; }

extern "C" const char* b2q_last_error(void) { return b2q::last_error_buffer(); }

extern "C" uint64_t b2q_launch_count(void) {
  return b2q::g_launch_count.load(std::memory_order_relaxed);
}

extern "C" int b2q_device_info(int* sm_count, uint64_t* hbm_bytes, int* cc) {
  int dev = 0;
  B2Q_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  B2Q_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (hbm_bytes) *hbm_bytes = (uint64_t)prop.totalGlobalMem;
  if (cc) *cc = prop.major * 10 + prop.minor;
  return B2Q_OK;
}

// ---- host-side scheduler helper ---------------------------------------------------------
// block <- (matrix on the row-index bits bitpos[]) . block, for the gate fuser
// (cirq_b200/fusion.py apply_to_block / expand_matrix): block is a 2^u x 2^u
// complex128 row-major matrix on the HOST, matrix a 2^k x 2^k complex128 whose
// first wire is its most significant index bit, bitpos[j] the bit of the block's
// row index that wire j acts on.  Small (u <= 6) and called once per circuit
// gate; numpy's tensordot + moveaxis cost ~10x more in dispatch than the
// arithmetic is worth.
extern "C" int b2q_host_left_apply(double* block, int u, const double* matrix, const int* bitpos,
                                   int k) {
  B2Q_REQUIRE(block != nullptr && matrix != nullptr && bitpos != nullptr, "null argument");
  B2Q_REQUIRE(u >= 1 && u <= 6 && k >= 1 && k <= u, "bad sizes u=%d k=%d", u, k);
  const int dim = 1 << u, d = 1 << k;
  int sorted[6];
  for (int j = 0; j < k; ++j) {
    B2Q_REQUIRE(bitpos[j] >= 0 && bitpos[j] < u, "bit position %d out of range", bitpos[j]);
    sorted[j] = bitpos[j];
  }
  for (int a = 1; a < k; ++a)
    for (int b = a; b > 0 && sorted[b - 1] > sorted[b]; --b) {
      const int t = sorted[b];
      sorted[b] = sorted[b - 1];
      sorted[b - 1] = t;
    }
  for (int a = 1; a < k; ++a) B2Q_REQUIRE(sorted[a] != sorted[a - 1], "duplicate bit position");
  int offset[64];
  for (int t = 0; t < d; ++t) {
    int o = 0;
    for (int j = 0; j < k; ++j)
      if ((t >> (k - 1 - j)) & 1) o |= 1 << bitpos[j];
    offset[t] = o;
  }
  double tmp[64 * 64 * 2];  // d rows of dim complex numbers
  const int groups = 1 << (u - k);
  for (int g = 0; g < groups; ++g) {
    const int base = (int)b2q::insert_zero_bits((uint64_t)g, sorted, k);
    for (int r = 0; r < d; ++r) {
      double* out = tmp + (size_t)r * dim * 2;
      for (int c = 0; c < 2 * dim; ++c) out[c] = 0.0;
      for (int t = 0; t < d; ++t) {
        const double mr = matrix[2 * (r * d + t)], mi = matrix[2 * (r * d + t) + 1];
        if (mr == 0.0 && mi == 0.0) continue;
        const double* row = block + (size_t)(base | offset[t]) * dim * 2;
        for (int c = 0; c < dim; ++c) {
          const double xr = row[2 * c], xi = row[2 * c + 1];
          out[2 * c] += mr * xr - mi * xi;
          out[2 * c + 1] += mr * xi + mi * xr;
        }
      }
    }
    for (int r = 0; r < d; ++r) {
      double* row = block + (size_t)(base | offset[r]) * dim * 2;
      const double* out = tmp + (size_t)r * dim * 2;
      for (int c = 0; c < 2 * dim; ++c) row[c] = out[c];
    }
  }
  return B2Q_OK;
}

// out <- M_{n-1} ... M_1 M_0 on a space of u wires: `out` (2^u x 2^u complex128, row-major)
// starts as the identity and is left-multiplied by every member in order; member m is a
// 2^k x 2^k matrix (ks[m] wires, consecutive in `matrices`) on the row-index bits
// bitpos[...] (consecutive in `bitpos`, first wire = most significant index bit).  The
// gate fuser keeps a block as the LIST of its member gates while it schedules and asks
// for the product once, when the block is emitted (one call per block instead of two
// to three per gate).
extern "C" int b2q_host_compose(double* out, int u, int num_members, const int* ks,
                                const int* bitpos, const double* matrices) {
  B2Q_REQUIRE(out != nullptr && ks != nullptr && bitpos != nullptr && matrices != nullptr, "null argument");
  B2Q_REQUIRE(u >= 1 && u <= 6 && num_members >= 1, "bad sizes u=%d members=%d", u, num_members);
  const int dim = 1 << u;
  for (int r = 0; r < dim; ++r)
    for (int c = 0; c < dim; ++c) {
      out[2 * ((size_t)r * dim + c)] = r == c ? 1.0 : 0.0;
      out[2 * ((size_t)r * dim + c) + 1] = 0.0;
    }
  size_t boff = 0, moff = 0;
  for (int m = 0; m < num_members; ++m) {
    const int k = ks[m];
    B2Q_REQUIRE(k >= 1 && k <= u, "member %d has %d wires", m, k);
    const int rc = b2q_host_left_apply(out, u, matrices + moff, bitpos + boff, k);
    if (rc != B2Q_OK) return rc;
    boff += k;
    moff += (size_t)2 << (2 * k);
  }
  return B2Q_OK;
}

// out[i] = prod_m diag_m[bits of i at member m's wires]: the table of a diagonal block
// that the fuser kept as a list of diagonal gates (u <= 16 wires; member m has ks[m]
// wires at index bits bitpos[...], first wire = most significant bit of its own index).
extern "C" int b2q_host_compose_diag(double* out, int u, int num_members, const int* ks,
                                     const int* bitpos, const double* diags) {
  B2Q_REQUIRE(out != nullptr && ks != nullptr && bitpos != nullptr && diags != nullptr, "null argument");
  B2Q_REQUIRE(u >= 1 && u <= 16 && num_members >= 1, "bad sizes u=%d members=%d", u, num_members);
  const int dim = 1 << u;
  for (int i = 0; i < dim; ++i) {
    out[2 * i] = 1.0;
    out[2 * i + 1] = 0.0;
  }
  size_t boff = 0, doff = 0;
  for (int m = 0; m < num_members; ++m) {
    const int k = ks[m];
    B2Q_REQUIRE(k >= 1 && k <= u, "member %d has %d wires", m, k);
    const int* pos = bitpos + boff;
    for (int q = 0; q < k; ++q) B2Q_REQUIRE(pos[q] >= 0 && pos[q] < u, "bit position out of range");
    const double* d = diags + doff;
    for (int i = 0; i < dim; ++i) {
      int sub = 0;
      for (int q = 0; q < k; ++q) sub = (sub << 1) | ((i >> pos[q]) & 1);
      const double dr = d[2 * sub], di = d[2 * sub + 1];
      const double xr = out[2 * i], xi = out[2 * i + 1];
      out[2 * i] = xr * dr - xi * di;
      out[2 * i + 1] = xr * di + xi * dr;
    }
    boff += k;
    doff += (size_t)2 << k;
  }
  return B2Q_OK;
}

// ---- a recorded schedule executed by ONE library call ---------------------------------
// (cirq_b200/program.py compiles the device operations of a cached schedule into this
// form; replaying them from Python costs ~20 us of interpreter + ctypes per operation,
// which is what a 20-qubit circuit's end-to-end time consists of.)

extern "C" int b2q_schedule_op_bytes(void) { return (int)sizeof(b2q_schedule_op); }

extern "C" int b2q_run_schedule(int dtype, int num_ops, const b2q_schedule_op* ops, const int* ints,
                                const double* reals, int num_slots, void* const* slots,
                                int* permute_passes, void* stream) {
  B2Q_REQUIRE(dtype == B2Q_C64 || dtype == B2Q_C128, "bad dtype %d", dtype);
  B2Q_REQUIRE(num_ops >= 0 && num_slots >= 0, "negative count");
  B2Q_REQUIRE(num_ops == 0 || (ops != nullptr && slots != nullptr), "null argument");
  for (int i = 0; i < num_ops; ++i) {
    const b2q_schedule_op& op = ops[i];
    B2Q_REQUIRE(op.slot >= 0 && op.slot < num_slots && slots[op.slot] != nullptr,
                "operation %d: bad state slot %d", i, op.slot);
    void* const st = slots[op.slot];
    const int* const iv = ints + op.ints_offset;
    const double* const rv = reals + op.reals_offset;
    int rc = B2Q_OK;
    switch (op.kind) {
      case B2Q_OP_BASIS:
        rc = b2q_sv_init_basis(st, dtype, op.n_bits, op.basis_index, stream);
        break;
      case B2Q_OP_KRON:
        B2Q_REQUIRE(op.a >= 0 && op.a < num_slots && op.b >= 0 && op.b < num_slots &&
                        slots[op.a] != nullptr && slots[op.b] != nullptr,
                    "operation %d: bad input slots", i);
        // ints: bits of a, bits of b
        rc = b2q_sv_kron(slots[op.a], iv[0], slots[op.b], iv[1], dtype, st, stream);
        break;
      case B2Q_OP_DENSE:  // ints: ks[count], then all targets; reals: the matrices
        rc = b2q_sv_apply_batch(st, dtype, op.n_bits, op.count, iv, iv + op.count, rv, nullptr, stream);
        break;
      case B2Q_OP_TILE:
        rc = b2q_sv_apply_tile_blocks(st, dtype, op.n_bits, op.count, iv, iv + op.count, rv, stream);
        break;
      case B2Q_OP_DIAGONAL:  // ints: targets[count]; reals: the 2^count entries
        rc = b2q_sv_apply_diagonal(st, dtype, op.n_bits, rv, iv, op.count, stream);
        break;
      case B2Q_OP_SCALE:
        rc = b2q_sv_scale(st, dtype, op.n_bits, rv[0], rv[1], stream);
        break;
      case B2Q_OP_PERMUTE: {
        int passes = 0;
        rc = b2q_sv_permute_bits_inplace(st, dtype, op.n_bits, iv, &passes, stream);
        if (permute_passes != nullptr) permute_passes[op.slot] += passes;
        break;
      }
      default:
        return b2q::set_error(B2Q_ERR_INVALID, "operation %d: unknown kind %d", i, op.kind);
    }
    if (rc != B2Q_OK) return rc;
  }
  return B2Q_OK;
}
