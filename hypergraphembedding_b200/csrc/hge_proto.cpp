// Wire-format reader / writer for the two messages on the path's boundary
// (hypergraph.proto:6-35): Hypergraph in, HypergraphEmbedding out -- host code.
//
// The reference walks the parsed proto in Python: ToCsrMatrix / ToEdgeCsrMatrix
// (hypergraph_util.py:96-135) append one list element per incidence, CompressRange / Relabel
// (:198-244) call AddNodeToEdge per incidence with a linear membership scan, and
// EmbedAlgebraicDistance packs its result one row at a time (algebraic_distance.py:169-174).
// Here the serialized bytes are read straight into id / row-pointer / column-id arrays, the
// compress + CSR + transpose step is done on those arrays, and the embedding is written as
// wire bytes that any protobuf runtime parses back (map entries in ascending key order, floats
// unpacked as proto2 serializers emit them).
//
//   map<int32, V> f = N   is   repeated Entry { int32 key = 1; V value = 2; }  on the wire.
//   Duplicate keys: the last entry wins, as in every protobuf runtime.
//   repeated int32 / float are accepted packed or unpacked.  Unknown fields are skipped.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "hge_common.cuh"

namespace {

struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;

  bool done() const { return p >= end; }

  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; shift <= 63; shift += 7) {
      if (p >= end) {
        ok = false;
        return 0;
      }
      const uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
    }
    ok = false;   // more than 10 bytes
    return 0;
  }

  Reader sub() {   // length-delimited payload
    const uint64_t n = varint();
    if (!ok || n > (uint64_t)(end - p)) {
      ok = false;
      return Reader{p, p};
    }
    Reader r{p, p + n};
    p += n;
    return r;
  }

  uint32_t fixed32() {
    if (end - p < 4) {
      ok = false;
      return 0;
    }
    uint32_t v;
    memcpy(&v, p, 4);
    p += 4;
    return v;
  }

  void skip(uint32_t wire_type) {
    switch (wire_type) {
      case 0: varint(); break;
      case 1: if (end - p < 8) ok = false; else p += 8; break;
      case 2: sub(); break;
      case 5: fixed32(); break;
      default: ok = false;   // groups are not used by this schema
    }
  }
};

// One side of a Hypergraph: the entries of `map<int32, NodeData> node` or `edge` in wire order.
struct Side {
  std::vector<int32_t> ids;
  std::vector<int64_t> ptr{0};
  std::vector<int32_t> members;
  std::vector<float> weight;
  std::vector<std::string> names;
  std::vector<uint8_t> has_name;
  bool any_name = false;
};

struct ParsedValue {
  std::vector<int32_t> members;
  float weight = 1.0f;   // [default = 1], hypergraph.proto:10,15
  std::string name;
  bool has_name = false;
};

bool parse_data(Reader r, ParsedValue* v) {   // NodeData / EdgeData
  while (r.ok && !r.done()) {
    const uint64_t tag = r.varint();
    const uint32_t field = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7);
    if (field == 1 && wt == 0) {
      v->members.push_back((int32_t)r.varint());
    } else if (field == 1 && wt == 2) {
      Reader packed = r.sub();
      while (packed.ok && !packed.done()) v->members.push_back((int32_t)packed.varint());
      r.ok = r.ok && packed.ok;
    } else if (field == 2 && wt == 2) {
      Reader s = r.sub();
      v->name.assign(reinterpret_cast<const char*>(s.p), (size_t)(s.end - s.p));
      v->has_name = true;
    } else if (field == 3 && wt == 5) {
      const uint32_t bits = r.fixed32();
      memcpy(&v->weight, &bits, 4);
    } else {
      r.skip(wt);
    }
  }
  return r.ok;
}

bool parse_entry(Reader r, int32_t* key, ParsedValue* v) {
  *key = 0;
  while (r.ok && !r.done()) {
    const uint64_t tag = r.varint();
    const uint32_t field = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7);
    if (field == 1 && wt == 0) {
      *key = (int32_t)r.varint();
    } else if (field == 2 && wt == 2) {
      if (!parse_data(r.sub(), v)) return false;
    } else {
      r.skip(wt);
    }
  }
  return r.ok;
}

// A key seen before does not add an entry: its value is parked in `later` and replaces the
// first occurrence after the scan ("last entry wins").
struct Replacements {
  std::unordered_map<int32_t, size_t> pos;   // key -> entry index
  std::unordered_map<size_t, ParsedValue> later;
};

void append_entry(Side* s, Replacements* rep, int32_t key, ParsedValue&& v) {
  auto it = rep->pos.find(key);
  if (it != rep->pos.end()) {
    rep->later[it->second] = std::move(v);
    return;
  }
  rep->pos[key] = s->ids.size();
  s->ids.push_back(key);
  s->members.insert(s->members.end(), v.members.begin(), v.members.end());
  s->ptr.push_back((int64_t)s->members.size());
  s->weight.push_back(v.weight);
  s->has_name.push_back(v.has_name ? 1 : 0);
  s->names.push_back(std::move(v.name));
  s->any_name |= v.has_name;
}

}  // namespace

struct hge_hypergraph {
  Side node, edge;
  std::string name;
  bool has_name = false;
};

struct hge_embedding {
  // entries of `map<int32, Embedding> node / edge` in wire order
  std::vector<int32_t> node_ids, edge_ids;
  std::vector<int64_t> node_ptr{0}, edge_ptr{0};
  std::vector<float> node_values, edge_values;
  int32_t dim = 0;
  bool has_dim = false;
  std::string method_name;
};

namespace {

// Applies "last entry wins" for keys that occurred more than once (rare: rebuilds the side).
void apply_replacements(Side* s, Replacements& rep) {
  if (rep.later.empty()) return;
  Side out;
  for (size_t i = 0; i < s->ids.size(); ++i) {
    auto it = rep.later.find(i);
    out.ids.push_back(s->ids[i]);
    if (it != rep.later.end()) {
      const ParsedValue& v = it->second;
      out.members.insert(out.members.end(), v.members.begin(), v.members.end());
      out.weight.push_back(v.weight);
      out.has_name.push_back(v.has_name ? 1 : 0);
      out.names.push_back(v.name);
      out.any_name |= v.has_name;
    } else {
      out.members.insert(out.members.end(), s->members.begin() + s->ptr[i],
                         s->members.begin() + s->ptr[i + 1]);
      out.weight.push_back(s->weight[i]);
      out.has_name.push_back(s->has_name[i]);
      out.names.push_back(s->names[i]);
      out.any_name |= s->has_name[i] != 0;
    }
    out.ptr.push_back((int64_t)out.members.size());
  }
  *s = std::move(out);
}

// ---- writer helpers ---------------------------------------------------------------------------
inline size_t varint_size(uint64_t v) {
  size_t n = 1;
  while (v >= 0x80) {
    v >>= 7;
    ++n;
  }
  return n;
}
inline uint8_t* put_varint(uint8_t* p, uint64_t v) {
  while (v >= 0x80) {
    *p++ = (uint8_t)(v | 0x80);
    v >>= 7;
  }
  *p++ = (uint8_t)v;
  return p;
}
inline uint64_t int32_as_varint(int32_t v) { return (uint64_t)(int64_t)v; }   // sign-extended

// Entry { key = 1, value = 2 }: both are always written, also a zero key
size_t embedding_entry_payload(int32_t key, int32_t R) {
  const size_t value_len = (size_t)R * 5;   // tag 0x0d + 4 bytes per float, unpacked
  return 1 + varint_size(int32_as_varint(key)) + 1 + varint_size(value_len) + value_len;
}

uint8_t* put_embedding_entry(uint8_t* p, uint8_t field_tag, int32_t key, const float* row, int32_t R) {
  const size_t value_len = (size_t)R * 5;
  *p++ = field_tag;
  p = put_varint(p, embedding_entry_payload(key, R));
  *p++ = 0x08;   // key, varint
  p = put_varint(p, int32_as_varint(key));
  *p++ = 0x12;   // value, length-delimited Embedding
  p = put_varint(p, value_len);
  for (int32_t c = 0; c < R; ++c) {
    *p++ = 0x0d;   // values = 1, fixed32
    memcpy(p, row + c, 4);
    p += 4;
  }
  return p;
}

}  // namespace

extern "C" {

int hge_hypergraph_parse(const void* buf, size_t len, hge_hypergraph** out) {
  HGE_REQUIRE(out && (buf || len == 0), "hge_hypergraph_parse: NULL argument");
  *out = nullptr;
  hge_hypergraph* hg = new (std::nothrow) hge_hypergraph();
  if (!hg) return HGE_ERR_NOMEM;
  Reader r{static_cast<const uint8_t*>(buf), static_cast<const uint8_t*>(buf) + len};
  Replacements node_rep, edge_rep;
  while (r.ok && !r.done()) {
    const uint64_t tag = r.varint();
    const uint32_t field = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7);
    if ((field == 1 || field == 2) && wt == 2) {
      int32_t key;
      ParsedValue v;
      if (!parse_entry(r.sub(), &key, &v)) {
        r.ok = false;
        break;
      }
      append_entry(field == 1 ? &hg->node : &hg->edge, field == 1 ? &node_rep : &edge_rep, key,
                   std::move(v));
    } else if (field == 3 && wt == 2) {
      Reader s = r.sub();
      hg->name.assign(reinterpret_cast<const char*>(s.p), (size_t)(s.end - s.p));
      hg->has_name = true;
    } else {
      r.skip(wt);
    }
  }
  if (!r.ok) {
    delete hg;
    hge_set_error("hge_hypergraph_parse: malformed or truncated Hypergraph message");
    return HGE_ERR_INVALID;
  }
  apply_replacements(&hg->node, node_rep);
  apply_replacements(&hg->edge, edge_rep);
  *out = hg;
  return HGE_OK;
}

int hge_hypergraph_destroy(hge_hypergraph* hg) {
  delete hg;
  return HGE_OK;
}

int hge_hypergraph_sizes(const hge_hypergraph* hg, int64_t* sizes4) {
  HGE_REQUIRE(hg && sizes4, "hge_hypergraph_sizes: NULL argument");
  sizes4[0] = (int64_t)hg->node.ids.size();
  sizes4[1] = (int64_t)hg->edge.ids.size();
  sizes4[2] = (int64_t)hg->node.members.size();
  sizes4[3] = (int64_t)hg->edge.members.size();
  return HGE_OK;
}

int hge_hypergraph_arrays(const hge_hypergraph* hg, int32_t* node_ids, int64_t* node_ptr,
                          int32_t* node_edges, float* node_weight, int32_t* edge_ids,
                          int64_t* edge_ptr, int32_t* edge_nodes, float* edge_weight) {
  HGE_REQUIRE(hg, "hge_hypergraph_arrays: NULL argument");
  auto put = [](const Side& s, int32_t* ids, int64_t* ptr, int32_t* members, float* weight) {
    if (ids) std::copy(s.ids.begin(), s.ids.end(), ids);
    if (ptr) std::copy(s.ptr.begin(), s.ptr.end(), ptr);
    if (members) std::copy(s.members.begin(), s.members.end(), members);
    if (weight) std::copy(s.weight.begin(), s.weight.end(), weight);
  };
  put(hg->node, node_ids, node_ptr, node_edges, node_weight);
  put(hg->edge, edge_ids, edge_ptr, edge_nodes, edge_weight);
  return HGE_OK;
}

int hge_hypergraph_compress(const hge_hypergraph* hg, int32_t* sorted_node_ids,
                            int32_t* sorted_edge_ids, int64_t* n2e_ptr, int32_t* n2e_idx,
                            int64_t* e2n_ptr, int32_t* e2n_idx, int64_t* nnz_out) {
  HGE_REQUIRE(hg && sorted_node_ids && sorted_edge_ids && n2e_ptr && n2e_idx && e2n_ptr && e2n_idx &&
                  nnz_out, "hge_hypergraph_compress: NULL argument");
  const Side& nd = hg->node;
  const Side& ed = hg->edge;
  const size_t N = nd.ids.size(), E = ed.ids.size();
  // sorted ids = the inverse maps of CompressRange (hypergraph_util.py:229-232)
  std::vector<int32_t> order(N);
  for (size_t i = 0; i < N; ++i) order[i] = (int32_t)i;
  std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return nd.ids[a] < nd.ids[b]; });
  for (size_t i = 0; i < N; ++i) sorted_node_ids[i] = nd.ids[(size_t)order[i]];
  std::copy(ed.ids.begin(), ed.ids.end(), sorted_edge_ids);
  std::sort(sorted_edge_ids, sorted_edge_ids + E);
  // rank of an original edge id: direct table when the id range is dense enough, else search
  const int64_t lo = E ? sorted_edge_ids[0] : 0, hi = E ? sorted_edge_ids[E - 1] : -1;
  const bool table = E > 0 && (hi - lo + 1) <= (int64_t)(8 * E + 1024);
  std::vector<int32_t> rank;
  if (table) {
    rank.assign((size_t)(hi - lo + 1), -1);
    for (size_t i = 0; i < E; ++i) rank[(size_t)(sorted_edge_ids[i] - lo)] = (int32_t)i;
  }
  auto edge_rank = [&](int32_t id) -> int32_t {
    if (table) return (id < lo || id > hi) ? -1 : rank[(size_t)(id - lo)];
    const int32_t* it = std::lower_bound(sorted_edge_ids, sorted_edge_ids + E, id);
    return (it != sorted_edge_ids + E && *it == id) ? (int32_t)(it - sorted_edge_ids) : -1;
  };
  // node rows in compressed order; column ids sorted, duplicates collapsed (ToCsrMatrix :96-114)
  int64_t nnz = 0;
  n2e_ptr[0] = 0;
  std::vector<int64_t> edge_count(E + 1, 0);
  for (size_t i = 0; i < N; ++i) {
    const size_t src = (size_t)order[i];
    const int64_t b = nd.ptr[src], e = nd.ptr[src + 1];
    int32_t* row = n2e_idx + nnz;
    int64_t m = 0;
    for (int64_t p = b; p < e; ++p) {
      const int32_t c = edge_rank(nd.members[(size_t)p]);
      if (c < 0) {   // Relabel asserts edge_idx in edge_map (hypergraph_util.py:207)
        hge_set_error("hge_hypergraph_compress: node %d lists edge %d, which is not a key of the "
                      "edge map", nd.ids[src], nd.members[(size_t)p]);
        return HGE_ERR_INVALID;
      }
      row[m++] = c;
    }
    std::sort(row, row + m);
    m = std::unique(row, row + m) - row;
    for (int64_t k = 0; k < m; ++k) edge_count[(size_t)row[k] + 1]++;
    nnz += m;
    n2e_ptr[i + 1] = nnz;
  }
  // transpose: after Relabel only node.edges drives the connections, so edge->node is A^T
  for (size_t j = 0; j < E; ++j) edge_count[j + 1] += edge_count[j];
  std::copy(edge_count.begin(), edge_count.end(), e2n_ptr);
  std::vector<int64_t> cursor(edge_count.begin(), edge_count.end() - 1);
  for (size_t i = 0; i < N; ++i)
    for (int64_t p = n2e_ptr[i]; p < n2e_ptr[i + 1]; ++p) e2n_idx[cursor[(size_t)n2e_idx[p]]++] = (int32_t)i;
  *nnz_out = nnz;
  return HGE_OK;
}

int hge_embedding_wire_size(const int32_t* node_ids, int64_t num_nodes, const int32_t* edge_ids,
                            int64_t num_edges, int32_t R, int32_t dim, const char* method_name,
                            size_t* out) {
  HGE_REQUIRE(out && R >= 0 && num_nodes >= 0 && num_edges >= 0 && (node_ids || num_nodes == 0) &&
                  (edge_ids || num_edges == 0), "hge_embedding_wire_size: bad argument");
  size_t total = 0;
  for (int64_t i = 0; i < num_nodes; ++i) {
    const size_t payload = embedding_entry_payload(node_ids[i], R);
    total += 1 + varint_size(payload) + payload;
  }
  for (int64_t i = 0; i < num_edges; ++i) {
    const size_t payload = embedding_entry_payload(edge_ids[i], R);
    total += 1 + varint_size(payload) + payload;
  }
  total += 1 + varint_size(int32_as_varint(dim));
  if (method_name) {
    const size_t n = strlen(method_name);
    total += 1 + varint_size(n) + n;
  }
  *out = total;
  return HGE_OK;
}

int hge_embedding_write(const int32_t* node_ids, int64_t num_nodes, const float* xn,
                        const int32_t* edge_ids, int64_t num_edges, const float* xe, int32_t R,
                        int32_t dim, const char* method_name, void* out, size_t capacity,
                        size_t* written) {
  size_t need = 0;
  HGE_TRY(hge_embedding_wire_size(node_ids, num_nodes, edge_ids, num_edges, R, dim, method_name, &need));
  HGE_REQUIRE(out && written && (xn || num_nodes == 0 || R == 0) && (xe || num_edges == 0 || R == 0),
              "hge_embedding_write: NULL argument");
  HGE_REQUIRE(capacity >= need, "hge_embedding_write: buffer of %zu bytes, %zu needed", capacity, need);
  for (int64_t i = 1; i < num_nodes; ++i)
    HGE_REQUIRE(node_ids[i - 1] < node_ids[i], "hge_embedding_write: node ids must ascend strictly");
  for (int64_t i = 1; i < num_edges; ++i)
    HGE_REQUIRE(edge_ids[i - 1] < edge_ids[i], "hge_embedding_write: edge ids must ascend strictly");
  uint8_t* p = static_cast<uint8_t*>(out);
  for (int64_t i = 0; i < num_nodes; ++i) p = put_embedding_entry(p, 0x0a, node_ids[i], xn + i * R, R);
  for (int64_t i = 0; i < num_edges; ++i) p = put_embedding_entry(p, 0x12, edge_ids[i], xe + i * R, R);
  *p++ = 0x18;   // dim = 3
  p = put_varint(p, int32_as_varint(dim));
  if (method_name) {
    const size_t n = strlen(method_name);
    *p++ = 0x22;   // method_name = 4
    p = put_varint(p, n);
    memcpy(p, method_name, n);
    p += n;
  }
  *written = (size_t)(p - static_cast<uint8_t*>(out));
  return HGE_OK;
}

int hge_embedding_parse(const void* buf, size_t len, hge_embedding** out) {
  HGE_REQUIRE(out && (buf || len == 0), "hge_embedding_parse: NULL argument");
  *out = nullptr;
  hge_embedding* emb = new (std::nothrow) hge_embedding();
  if (!emb) return HGE_ERR_NOMEM;
  Reader r{static_cast<const uint8_t*>(buf), static_cast<const uint8_t*>(buf) + len};
  std::unordered_map<int32_t, size_t> seen[2];
  bool duplicate = false;
  while (r.ok && !r.done()) {
    const uint64_t tag = r.varint();
    const uint32_t field = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7);
    if ((field == 1 || field == 2) && wt == 2) {
      std::vector<int32_t>& ids = field == 1 ? emb->node_ids : emb->edge_ids;
      std::vector<int64_t>& ptr = field == 1 ? emb->node_ptr : emb->edge_ptr;
      std::vector<float>& vals = field == 1 ? emb->node_values : emb->edge_values;
      Reader e = r.sub();
      int32_t key = 0;
      while (e.ok && !e.done()) {
        const uint64_t t = e.varint();
        const uint32_t f = (uint32_t)(t >> 3), w = (uint32_t)(t & 7);
        if (f == 1 && w == 0) {
          key = (int32_t)e.varint();
        } else if (f == 2 && w == 2) {
          Reader v = e.sub();
          while (v.ok && !v.done()) {
            const uint64_t vt = v.varint();
            const uint32_t vf = (uint32_t)(vt >> 3), vw = (uint32_t)(vt & 7);
            if (vf == 1 && vw == 5) {
              const uint32_t bits = v.fixed32();
              float x;
              memcpy(&x, &bits, 4);
              vals.push_back(x);
            } else if (vf == 1 && vw == 2) {
              Reader packed = v.sub();
              if ((packed.end - packed.p) % 4 != 0) v.ok = false;
              while (v.ok && !packed.done()) {
                const uint32_t bits = packed.fixed32();
                float x;
                memcpy(&x, &bits, 4);
                vals.push_back(x);
              }
            } else {
              v.skip(vw);
            }
          }
          e.ok = e.ok && v.ok;
        } else {
          e.skip(w);
        }
      }
      r.ok = r.ok && e.ok;
      duplicate |= !seen[field - 1].emplace(key, ids.size()).second;
      ids.push_back(key);
      ptr.push_back((int64_t)vals.size());
    } else if (field == 3 && wt == 0) {
      emb->dim = (int32_t)r.varint();
      emb->has_dim = true;
    } else if (field == 4 && wt == 2) {
      Reader s = r.sub();
      emb->method_name.assign(reinterpret_cast<const char*>(s.p), (size_t)(s.end - s.p));
    } else {
      r.skip(wt);
    }
  }
  if (!r.ok || duplicate) {
    delete emb;
    hge_set_error(duplicate ? "hge_embedding_parse: a map key occurs twice"
                            : "hge_embedding_parse: malformed or truncated HypergraphEmbedding message");
    return duplicate ? HGE_ERR_UNSUPPORTED : HGE_ERR_INVALID;
  }
  *out = emb;
  return HGE_OK;
}

int hge_embedding_destroy(hge_embedding* emb) {
  delete emb;
  return HGE_OK;
}

int hge_embedding_sizes(const hge_embedding* emb, int64_t* sizes4, int32_t* dim) {
  HGE_REQUIRE(emb && sizes4, "hge_embedding_sizes: NULL argument");
  sizes4[0] = (int64_t)emb->node_ids.size();
  sizes4[1] = (int64_t)emb->edge_ids.size();
  sizes4[2] = (int64_t)emb->node_values.size();
  sizes4[3] = (int64_t)emb->edge_values.size();
  if (dim) *dim = emb->has_dim ? emb->dim : -1;
  return HGE_OK;
}

int hge_embedding_arrays(const hge_embedding* emb, int32_t* node_ids, int64_t* node_ptr,
                         float* node_values, int32_t* edge_ids, int64_t* edge_ptr,
                         float* edge_values) {
  HGE_REQUIRE(emb, "hge_embedding_arrays: NULL argument");
  if (node_ids) std::copy(emb->node_ids.begin(), emb->node_ids.end(), node_ids);
  if (node_ptr) std::copy(emb->node_ptr.begin(), emb->node_ptr.end(), node_ptr);
  if (node_values) std::copy(emb->node_values.begin(), emb->node_values.end(), node_values);
  if (edge_ids) std::copy(emb->edge_ids.begin(), emb->edge_ids.end(), edge_ids);
  if (edge_ptr) std::copy(emb->edge_ptr.begin(), emb->edge_ptr.end(), edge_ptr);
  if (edge_values) std::copy(emb->edge_values.begin(), emb->edge_values.end(), edge_values);
  return HGE_OK;
}

}  // extern "C"
