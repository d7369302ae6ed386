// circuit.cuh -- levelised boolean circuits behind the C ABI.  Included at the end of
// engine.cu (it uses the engine's internals: run_device, key_switch, the scratch allocators).
//
// The reference evaluates circuits one gate at a time (examples/add_two_numbers.rs:11-49: a
// ripple-carry adder of xor/and/or; src/gates.rs:157-199: mux; src/circuits.rs: compare_bit).  Here a
// circuit is recorded once over integer wire ids, split into levels of mutually independent
// bootstraps, and every level runs as ONE device batch over all `batch` independent input sets:
//   gather kernel   : wire pairs -> the contiguous in_pairs image of the batch (NOT = exact negation
//                     and constants are resolved on the fly: they never occupy a wire)
//   K0+K3           : one mixed-gate blind rotation for the whole level
//   mux combine     : the sound fused MUX adds its two level-1 samples (+1/8) before the key switch
//   K4              : one key switch, written straight into the level's (contiguous) wire slots
// Wires stay resident in HBM between levels; nothing but inputs and outputs crosses PCIe.
//
// Fused MUX (sel ? a : b): u1 = AND(sel, a) and u2 = ANDNY(sel, b) = (!sel AND b) are blind-rotated and
// extracted at level 1 with the CORRECT ring degree (trlwe.rs:106-120 -- the reference's optimised
// Gates::mux, gates.rs:157-183, extracts with N = 700 index arithmetic and is unsound),
// u1 + u2 + (0, 1/8) has phase +-1/8 exactly like an OR of two gate outputs, and ONE key switch brings
// it to level 0: two blind rotations + one key switch instead of mux_naive's three bootstraps.

namespace {

struct CirSrc {            // one operand of a bootstrapped gate / one output
  uint32_t wire;           // physical wire (kind 0), or unused
  uint8_t kind;            // 0 wire, 1 constant false, 2 constant true
  uint8_t neg;             // exact negation (gates.rs:202-204)
  uint16_t pad;
};

// in_pairs[(entry * batch + b)][2][w] from the wire store [wire][batch][w]
__global__ void circuit_gather_kernel(const uint32_t *__restrict__ wires, const CirSrc *__restrict__ src,
                                      uint32_t *__restrict__ dst, uint32_t operands, size_t batch, uint32_t w) {
  const size_t row = blockIdx.x;                         // (entry * operands + operand) * batch + b ... see below
  const size_t ent_op = row / batch, b = row % batch;
  const CirSrc s = src[ent_op];
  // destination: entry-major, then batch, then operand: [(entry * batch + b) * operands + operand][w]
  const size_t entry = ent_op / operands, opnd = ent_op % operands;
  uint32_t *d = dst + ((entry * batch + b) * operands + opnd) * w;
  if (s.kind == 0) {
    const uint32_t *p = wires + ((size_t)s.wire * batch + b) * w;
    for (uint32_t x = threadIdx.x; x < w; x += blockDim.x) d[x] = s.neg ? 0u - p[x] : p[x];
  } else {
    // gates.rs:212-218: constant(true) = (0, mu), constant(false) = (0, 1 - mu) [sic, release-mode wrap]
    const uint32_t mu = 0x20000000u;
    uint32_t v = s.kind == 2 ? mu : 1u - mu;
    if (s.neg) v = 0u - v;
    for (uint32_t x = threadIdx.x; x < w; x += blockDim.x) d[x] = x == w - 1 ? v : 0u;
  }
}

// ext[x] = u1 + u2 + (0, ..., 0, 1/8): the OR pre-combination (gates.rs:62-66) at level 1
__global__ void circuit_mux_combine_kernel(uint32_t *__restrict__ u1, const uint32_t *__restrict__ u2,
                                           size_t rows) {
  const uint32_t W1 = TFHE_N + 1;
  const size_t total = rows * W1;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
    uint32_t v = u1[x] + u2[x];
    if (x % W1 == TFHE_N) v += 0x20000000u;
    u1[x] = v;
  }
}

enum { CIR_INPUT = 0, CIR_CONST0, CIR_CONST1, CIR_NOT, CIR_GATE, CIR_MUX };
struct CirNode {
  int kind;
  int op;                 // tfhe_gate for CIR_GATE
  uint32_t a, b, c;       // operands (node ids); MUX: a = sel, b = then, c = else
  uint32_t depth;
  uint32_t phys;          // physical wire of INPUT / GATE / MUX nodes
};

}  // namespace

struct tfhe_circuit {
  tfhe_engine *e = nullptr;
  int dev = 0;              // the engine's device: destroy must not touch an engine that may be gone already
  std::vector<CirNode> nodes;
  std::vector<uint32_t> inputs, outputs;
  // compiled form
  bool compiled = false;
  uint32_t n_phys = 0, n_levels = 0, n_pbs = 0;
  struct Level { std::vector<CirSrc> src; std::vector<uint8_t> ops; uint32_t gates = 0, muxes = 0, phys0 = 0; };
  std::vector<Level> levels;
  std::vector<CirSrc> out_src;
  Scratch d_wires, d_pairs, d_ext, d_src, d_ops, d_out;
  // device-resident schedule: every level's operand list (then the outputs') in d_src, uploaded once per compile;
  // the per-row gate codes in d_ops, rebuilt only when the batch size changes
  std::vector<size_t> src_off, ops_off;     // per level (+ one entry for the outputs in src_off)
  bool src_on_device = false;
  size_t ops_batch = 0;
  std::vector<cudaEvent_t> ev;              // 4 per level: blind rotation start / end, key switch start / end
};

namespace {

CirSrc cir_resolve(const tfhe_circuit *c, uint32_t id) {
  CirSrc s{0, 0, 0, 0};
  for (;;) {
    const CirNode &n = c->nodes[id];
    if (n.kind == CIR_NOT) { s.neg ^= 1; id = n.a; continue; }
    if (n.kind == CIR_CONST0 || n.kind == CIR_CONST1) { s.kind = n.kind == CIR_CONST1 ? 2 : 1; return s; }
    s.wire = n.phys;
    return s;
  }
}

int cir_compile(tfhe_circuit *c) {
  if (c->compiled) return TFHE_OK;
  uint32_t maxd = 0;
  for (CirNode &n : c->nodes) {
    switch (n.kind) {
      case CIR_INPUT: case CIR_CONST0: case CIR_CONST1: n.depth = 0; break;
      case CIR_NOT: n.depth = c->nodes[n.a].depth; break;
      case CIR_GATE: n.depth = 1 + std::max(c->nodes[n.a].depth, c->nodes[n.b].depth); break;
      default: n.depth = 1 + std::max(c->nodes[n.a].depth, std::max(c->nodes[n.b].depth, c->nodes[n.c].depth));
    }
    maxd = std::max(maxd, n.depth);
  }
  // physical wires: inputs first, then level by level (binary gates, then muxes) so that every
  // level's key-switch output is one contiguous block of the wire store
  uint32_t phys = 0;
  for (uint32_t id : c->inputs) c->nodes[id].phys = phys++;
  c->levels.assign(maxd, tfhe_circuit::Level());
  c->n_pbs = 0;
  for (uint32_t d = 1; d <= maxd; d++) {
    tfhe_circuit::Level &lv = c->levels[d - 1];
    lv.phys0 = phys;
    for (CirNode &n : c->nodes) if (n.kind == CIR_GATE && n.depth == d) { n.phys = phys++; lv.gates++; }
    for (CirNode &n : c->nodes) if (n.kind == CIR_MUX && n.depth == d) { n.phys = phys++; lv.muxes++; }
  }
  for (uint32_t d = 1; d <= maxd; d++) {   // operands (all wires they name are now placed)
    tfhe_circuit::Level &lv = c->levels[d - 1];
    for (const CirNode &n : c->nodes)
      if (n.kind == CIR_GATE && n.depth == d) {
        lv.src.push_back(cir_resolve(c, n.a)); lv.src.push_back(cir_resolve(c, n.b));
        lv.ops.push_back((uint8_t)n.op);
      }
    // mux halves: all AND(sel, then) first, then all ANDNY(sel, else)
    for (const CirNode &n : c->nodes)
      if (n.kind == CIR_MUX && n.depth == d) {
        lv.src.push_back(cir_resolve(c, n.a)); lv.src.push_back(cir_resolve(c, n.b));
        lv.ops.push_back((uint8_t)TFHE_GATE_AND);
      }
    for (const CirNode &n : c->nodes)
      if (n.kind == CIR_MUX && n.depth == d) {
        lv.src.push_back(cir_resolve(c, n.a)); lv.src.push_back(cir_resolve(c, n.c));
        lv.ops.push_back((uint8_t)TFHE_GATE_ANDNY);
      }
    c->n_pbs += lv.gates + 2 * lv.muxes;
  }
  c->out_src.clear();
  for (uint32_t id : c->outputs) c->out_src.push_back(cir_resolve(c, id));
  c->n_phys = phys;
  c->n_levels = maxd;
  c->compiled = true;
  c->src_on_device = false;
  c->ops_batch = 0;
  return TFHE_OK;
}

int cir_new_node(tfhe_circuit *c, CirNode n, uint32_t *wire) {
  if (!c || !wire) return fail(TFHE_ERR_INVALID, "null argument");
  const uint32_t cnt = (uint32_t)c->nodes.size();
  if ((n.kind == CIR_NOT || n.kind == CIR_GATE || n.kind == CIR_MUX) && n.a >= cnt) return fail(TFHE_ERR_INVALID, "unknown wire %u", n.a);
  if ((n.kind == CIR_GATE || n.kind == CIR_MUX) && n.b >= cnt) return fail(TFHE_ERR_INVALID, "unknown wire %u", n.b);
  if (n.kind == CIR_MUX && n.c >= cnt) return fail(TFHE_ERR_INVALID, "unknown wire %u", n.c);
  c->nodes.push_back(n);
  c->compiled = false;
  *wire = cnt;
  return TFHE_OK;
}

}  // namespace

extern "C" {

int tfhe_circuit_create(tfhe_engine *e, tfhe_circuit **out) {
  if (!e || !out) return fail(TFHE_ERR_INVALID, "null argument");
  if (!e->peers.empty()) return fail(TFHE_ERR_INVALID, "circuits run on a one-GPU engine (levels are dependent; shard the batch over engines)");
  tfhe_circuit *c = new (std::nothrow) tfhe_circuit();
  if (!c) return fail(TFHE_ERR_ALLOC, "out of host memory");
  c->e = e;
  c->dev = e->dev;
  *out = c;
  return TFHE_OK;
}

void tfhe_circuit_destroy(tfhe_circuit *c) {
  if (!c) return;
  cudaSetDevice(c->dev);
  for (Scratch *s : {&c->d_wires, &c->d_pairs, &c->d_ext, &c->d_src, &c->d_ops, &c->d_out}) s->release();
  for (cudaEvent_t ev : c->ev) cudaEventDestroy(ev);
  delete c;
}

int tfhe_circuit_input(tfhe_circuit *c, uint32_t *wire) {
  int rc = cir_new_node(c, CirNode{CIR_INPUT, 0, 0, 0, 0, 0, 0}, wire);
  if (rc == TFHE_OK) c->inputs.push_back(*wire);
  return rc;
}
int tfhe_circuit_constant(tfhe_circuit *c, int value, uint32_t *wire) {
  return cir_new_node(c, CirNode{value ? CIR_CONST1 : CIR_CONST0, 0, 0, 0, 0, 0, 0}, wire);
}
int tfhe_circuit_not(tfhe_circuit *c, uint32_t a, uint32_t *wire) {
  return cir_new_node(c, CirNode{CIR_NOT, 0, a, 0, 0, 0, 0}, wire);
}
int tfhe_circuit_gate(tfhe_circuit *c, tfhe_gate op, uint32_t a, uint32_t b, uint32_t *wire) {
  if ((int)op < 0 || (int)op >= TFHE_GATE_COUNT) return fail(TFHE_ERR_INVALID, "bad gate %d", (int)op);
  return cir_new_node(c, CirNode{CIR_GATE, (int)op, a, b, 0, 0, 0}, wire);
}
int tfhe_circuit_mux(tfhe_circuit *c, uint32_t sel, uint32_t then_w, uint32_t else_w, uint32_t *wire) {
  return cir_new_node(c, CirNode{CIR_MUX, 0, sel, then_w, else_w, 0, 0}, wire);
}
int tfhe_circuit_output(tfhe_circuit *c, uint32_t wire) {
  if (!c) return fail(TFHE_ERR_INVALID, "null circuit");
  if (wire >= c->nodes.size()) return fail(TFHE_ERR_INVALID, "unknown wire %u", wire);
  c->outputs.push_back(wire);
  c->compiled = false;
  return TFHE_OK;
}
int tfhe_circuit_stats(tfhe_circuit *c, uint32_t *levels, uint32_t *bootstraps, uint32_t *key_switches) {
  if (!c) return fail(TFHE_ERR_INVALID, "null circuit");
  int rc = cir_compile(c);
  if (rc != TFHE_OK) return rc;
  uint32_t ks = 0;
  for (const auto &lv : c->levels) ks += lv.gates + lv.muxes;
  if (levels) *levels = c->n_levels;
  if (bootstraps) *bootstraps = c->n_pbs;
  if (key_switches) *key_switches = ks;
  return TFHE_OK;
}

int tfhe_circuit_run(tfhe_circuit *c, const uint32_t *inputs, uint32_t *outputs, size_t batch) {
  if (!c) return fail(TFHE_ERR_INVALID, "null circuit");
  if (batch == 0) return TFHE_OK;
  if ((!inputs && !c->inputs.empty()) || !outputs) return fail(TFHE_ERR_INVALID, "null buffer");
  int rc = cir_compile(c);
  if (rc != TFHE_OK) return rc;
  tfhe_engine *e = c->e;
  std::lock_guard<std::mutex> lock(e->mu);
  CU(cudaSetDevice(e->dev));
  if (!e->key_loaded) return fail(TFHE_ERR_NO_KEY, "cloud key not loaded");
  const uint32_t w = e->p.n + 1;
  const size_t W1 = TFHE_N + 1;
  CU(c->d_wires.reserve((size_t)std::max(c->n_phys, 1u) * batch * w * 4));
  uint32_t *wires = static_cast<uint32_t *>(c->d_wires.p);
  CU(cudaMemcpyAsync(wires, inputs, c->inputs.size() * batch * w * 4, cudaMemcpyHostToDevice, e->stream));
  float br_ms = 0.f, ks_ms = 0.f;
  tfhe_engine::Slot &sl = e->slot[0];
  // the schedule goes to the device once: operand lists at the first run after a compile, gate codes per batch size;
  // a run then queues gather / blind rotation / combine / key switch level after level WITHOUT touching the host
  // in between (kernel times come from per-level events read after the last level)
  if (!c->src_on_device) {
    std::vector<CirSrc> all;
    c->src_off.clear();
    for (const tfhe_circuit::Level &lv : c->levels) { c->src_off.push_back(all.size()); all.insert(all.end(), lv.src.begin(), lv.src.end()); }
    c->src_off.push_back(all.size());
    all.insert(all.end(), c->out_src.begin(), c->out_src.end());
    CU(c->d_src.reserve(std::max<size_t>(all.size(), 1) * sizeof(CirSrc)));
    CU(cudaMemcpyAsync(c->d_src.p, all.data(), all.size() * sizeof(CirSrc), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));   // `all` goes out of scope
    c->src_on_device = true;
  }
  if (c->ops_batch != batch) {
    std::vector<uint8_t> ops;
    c->ops_off.clear();
    for (const tfhe_circuit::Level &lv : c->levels) {
      const size_t entries = lv.gates + 2 * (size_t)lv.muxes;
      if (entries * batch > (size_t)1 << 22) return fail(TFHE_ERR_INVALID, "level too wide (%zu bootstraps); split the batch", entries * batch);
      c->ops_off.push_back(ops.size());
      for (size_t en = 0; en < entries; en++) ops.insert(ops.end(), batch, lv.ops[en]);
    }
    CU(c->d_ops.reserve(std::max<size_t>(ops.size(), 1)));
    CU(cudaMemcpyAsync(c->d_ops.p, ops.data(), ops.size(), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    c->ops_batch = batch;
  }
  while (c->ev.size() < 4 * c->levels.size()) {
    cudaEvent_t ev;
    CU(cudaEventCreate(&ev));
    c->ev.push_back(ev);
  }
  {   // size the level buffers once: a reallocation between levels would synchronise the device
    size_t max_rows = 0;
    for (const tfhe_circuit::Level &lv : c->levels) max_rows = std::max(max_rows, (lv.gates + 2 * (size_t)lv.muxes) * batch);
    CU(c->d_pairs.reserve(std::max<size_t>(max_rows, 1) * 2 * w * 4));
    CU(c->d_ext.reserve(std::max<size_t>(max_rows, 1) * W1 * 4));
    CU(c->d_out.reserve(std::max<size_t>(c->out_src.size() * batch, 1) * w * 4));
  }
  const CirSrc *d_src = static_cast<const CirSrc *>(c->d_src.p);
  for (size_t li = 0; li < c->levels.size(); li++) {
    const tfhe_circuit::Level &lv = c->levels[li];
    const size_t entries = lv.gates + 2 * (size_t)lv.muxes, rows = entries * batch;
    const size_t ks_rows = (lv.gates + (size_t)lv.muxes) * batch;
    circuit_gather_kernel<<<(unsigned)(rows * 2), 128, 0, e->stream>>>(
        wires, d_src + c->src_off[li], static_cast<uint32_t *>(c->d_pairs.p), 2, batch, w);
    CU(cudaGetLastError());
    e->launches++;
    uint32_t *ext = static_cast<uint32_t *>(c->d_ext.p);
    CU(cudaEventRecord(c->ev[4 * li + 0], e->stream));
    for (size_t base = 0; base < rows; base += kChunk) {   // blind rotation, extracted at level 1
      const size_t n = rows - base < kChunk ? rows - base : kChunk;
      rc = run_device(e, sl, 0, static_cast<const uint8_t *>(c->d_ops.p) + c->ops_off[li] + base,
                      -1, static_cast<const uint32_t *>(c->d_pairs.p) + base * 2 * w, ext + base * W1, n, 3);
      if (rc != TFHE_OK) return rc;
    }
    CU(cudaEventRecord(c->ev[4 * li + 1], e->stream));
    if (lv.muxes) {
      const size_t mrows = (size_t)lv.muxes * batch;
      uint32_t *u1 = ext + (size_t)lv.gates * batch * W1;
      circuit_mux_combine_kernel<<<(unsigned)std::min<size_t>((mrows * W1 + 255) / 256, 4096), 256, 0, e->stream>>>(
          u1, u1 + mrows * W1, mrows);
      CU(cudaGetLastError());
      e->launches++;
    }
    CU(cudaEventRecord(c->ev[4 * li + 2], e->stream));
    uint32_t *dst = wires + (size_t)lv.phys0 * batch * w;
    for (size_t base = 0; base < ks_rows; base += kChunk) {
      const size_t n = ks_rows - base < kChunk ? ks_rows - base : kChunk;
      rc = key_switch(e, ext + base * W1, dst + base * w, n);
      if (rc != TFHE_OK) return rc;
    }
    CU(cudaEventRecord(c->ev[4 * li + 3], e->stream));
  }
  // outputs (a NOT / constant / input may be an output): gather into [n_out][batch][w], then D2H
  if (!c->out_src.empty()) {
    const size_t rows = c->out_src.size() * batch;
    circuit_gather_kernel<<<(unsigned)rows, 128, 0, e->stream>>>(
        wires, d_src + c->src_off[c->levels.size()], static_cast<uint32_t *>(c->d_out.p), 1, batch, w);
    CU(cudaGetLastError());
    e->launches++;
    CU(cudaMemcpyAsync(outputs, c->d_out.p, rows * w * 4, cudaMemcpyDeviceToHost, e->stream));
  }
  CU(cudaStreamSynchronize(e->stream));
  for (size_t li = 0; li < c->levels.size(); li++) {
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, c->ev[4 * li + 0], c->ev[4 * li + 1]));
    br_ms += t;
    CU(cudaEventElapsedTime(&t, c->ev[4 * li + 2], c->ev[4 * li + 3]));
    ks_ms += t;
  }
  CU(cudaStreamSynchronize(e->stream));
  e->last_ms[0] = br_ms; e->last_ms[1] = ks_ms;
  return TFHE_OK;
}

}  // extern "C"
