// Plan compiler: integer-only, host-side.  Turns the reference's cluster-graph
// output (belief dimensions, scopeindex maps, spanning-tree edge lists) into
// HBM slot layouts, deduplicated gather/scatter tables and launch steps.
#include <algorithm>
#include <mutex>
#include <set>

#include "pgbp_internal.h"
#include "pgbp_shapes.h"

namespace pgbp {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const std::string& last_error() { return g_err; }

}  // namespace pgbp

using namespace pgbp;

int32_t pgbp_plan::intern_table(const std::vector<int32_t>& t) {
  auto it = tab_index.find(t);
  if (it != tab_index.end()) return it->second;
  int32_t off = (int32_t)tab.size();
  tab.insert(tab.end(), t.begin(), t.end());
  tab_index.emplace(t, off);
  return off;
}

int32_t pgbp_plan::find_table(const std::vector<int32_t>& t) const {
  auto it = tab_index.find(t);
  return it == tab_index.end() ? -1 : it->second;
}

// gather table of the sender in [I;K] order and scatter table of the receiver for message from -> to through j
static void msg_tables(const pgbp_plan* p, int mF, int s, const std::vector<int32_t>& upF, const std::vector<int32_t>& upT,
                       std::vector<int32_t>* g, std::vector<int32_t>* sc) {
  (void)p;
  std::vector<char> keep(mF, 0);
  for (int k : upF) keep[k] = 1;
  std::vector<int32_t> perm;
  perm.reserve(mF);
  for (int k = 0; k < mF; k++) if (!keep[k]) perm.push_back(k);
  for (int k : upF) perm.push_back(k);
  g->clear();
  g->reserve(tri(mF) + mF);
  for (int c = 0; c < mF; c++)
    for (int r = 0; r <= c; r++) {
      int a = perm[r], b = perm[c];
      g->push_back(pk(std::min(a, b), std::max(a, b)));
    }
  for (int k = 0; k < mF; k++) g->push_back(perm[k]);
  sc->clear();
  sc->reserve(tri(s) + s);
  for (int c = 0; c < s; c++)
    for (int r = 0; r <= c; r++) sc->push_back(pk(upT[r], upT[c]));
  for (int k = 0; k < s; k++) sc->push_back(upT[k]);
}

// The plan is immutable once created (it is shared by every batch, possibly from several host threads): the
// tables of BOTH directed messages of EVERY sepset are interned by pgbp_plan_create, so a descriptor can be
// built for any edge (pgbp_propagate on an edge that is in no tree) without growing `tab`.
int pgbp_plan::make_msg(int32_t from, int32_t j, int32_t to, MsgDesc* out) const {
  if (j < 0 || j >= nsepsets) PGBP_FAIL(PGBP_EINVAL, "sepset %d out of range", j);
  const std::vector<int32_t>*upF, *upT;
  int side;
  if (from == sep_a[j] && to == sep_b[j]) {
    upF = &up_a[j]; upT = &up_b[j]; side = 1;
  } else if (from == sep_b[j] && to == sep_a[j]) {
    upF = &up_b[j]; upT = &up_a[j]; side = 0;
  } else {
    PGBP_FAIL(PGBP_EINVAL, "sepset %d does not join clusters %d and %d", j, from, to);
  }
  const int mF = dim[from], s = dim[nclusters + j];
  MsgDesc m;
  m.fJ = jslot[from]; m.fh = hslot[from]; m.fg = gslot[from];
  m.sJ = jslot[nclusters + j]; m.sh = hslot[nclusters + j]; m.sg = gslot[nclusters + j];
  m.tJ = jslot[to]; m.th = hslot[to]; m.tg = gslot[to];
  m.dmsg = 2 * j + side;
  m.rJ = rjslot[m.dmsg]; m.rh = rhslot[m.dmsg];
  m.mF = mF; m.s = s; m.ref = 0;
  m.wid = walk_shape_id(ntraits, mF - s, s);
  std::vector<int32_t> g, sc;
  msg_tables(this, mF, s, *upF, *upT, &g, &sc);
  m.gat = find_table(g);
  m.sca = find_table(sc);
  if (m.gat < 0 || m.sca < 0) PGBP_FAIL(PGBP_ESTATE, "internal: index tables of sepset %d missing from the plan", j);
  *out = m;
  return 0;
}

static void msg_cost(const MsgDesc& m, double* b0, double* b1, double* fl) {
  const double S_F = tri(m.mF), S_s = tri(m.s), s = m.s, i = m.mF - m.s;
  const double base = (S_F + m.mF + 1) + 4 * (S_s + s + 1);
  *b0 += 8 * base;
  *b1 += 8 * (base + S_s + s);
  *fl += i * i * i / 3 + s * i * i + s * s * i + 2 * i * i + 2 * s * i + 4 * s * s;
}

// Cut a message sequence (reference order) into launch steps: messages of one
// step touch disjoint beliefs; per-belief read/write order is the reference's.
static int build_traversal(pgbp_plan* p, const std::vector<int32_t>& from, const std::vector<int32_t>& sep,
                           const std::vector<int32_t>& to, Traversal* tv) {
  const int n = (int)from.size();
  std::vector<int32_t> last_write(p->nbeliefs, -1), last_read(p->nbeliefs, -1);
  std::vector<MsgDesc> msgs(n);
  std::vector<int32_t> step(n);
  int nsteps = 0;
  for (int r = 0; r < n; r++) {
    PGBP_TRY(p->make_msg(from[r], sep[r], to[r], &msgs[r]));
    msgs[r].ref = r;
    const int f = from[r], t = to[r], sb = p->nclusters + sep[r];
    int st = 0;
    st = std::max(st, last_write[f] + 1);                       // RAW on the sender
    st = std::max(st, std::max(last_write[t], last_read[t]) + 1);   // WAW / WAR on the receiver
    st = std::max(st, std::max(last_write[sb], last_read[sb]) + 1); // sepset
    step[r] = st;
    last_read[f] = std::max(last_read[f], st);
    last_write[t] = st;
    last_write[sb] = st;
    nsteps = std::max(nsteps, st + 1);
    msg_cost(msgs[r], &tv->bytes_noresid, &tv->bytes_resid, &tv->flops);
  }
  // execution order: by step, then shape class, then reference order
  struct Key { int step, ci, cs, maxm, ref; };
  std::vector<Key> keys(n);
  for (int r = 0; r < n; r++) {
    int i = msgs[r].mF - msgs[r].s, s = msgs[r].s, ci, cs, maxm = 0;
    shape_class(i, s, &ci, &cs, &maxm);
    keys[r] = {step[r], ci, cs, maxm, r};
  }
  std::vector<int> ord(n);
  for (int r = 0; r < n; r++) ord[r] = r;
  std::sort(ord.begin(), ord.end(), [&](int a, int b) {
    const Key &x = keys[a], &y = keys[b];
    if (x.step != y.step) return x.step < y.step;
    if (x.ci != y.ci) return x.ci < y.ci;
    if (x.cs != y.cs) return x.cs < y.cs;
    if (x.maxm != y.maxm) return x.maxm < y.maxm;
    return x.ref < y.ref;
  });
  tv->msgs.resize(n);
  tv->step_of_msg.resize(n);
  tv->groups.clear();
  for (int k = 0; k < n; k++) {
    const Key& key = keys[ord[k]];
    tv->msgs[k] = msgs[ord[k]];
    tv->step_of_msg[k] = key.step;
    if (tv->groups.empty() || tv->groups.back().step != key.step || tv->groups.back().ci != key.ci ||
        tv->groups.back().cs != key.cs || tv->groups.back().maxm != key.maxm) {
      tv->groups.push_back({key.step, key.ci, key.cs, key.maxm, k, 0});
    }
    tv->groups.back().count++;
  }
  tv->nsteps = nsteps;
  tv->step_off.assign(nsteps + 1, 0);
  for (int k = 0; k < n; k++) tv->step_off[tv->step_of_msg[k] + 1]++;
  for (int st = 0; st < nsteps; st++) tv->step_off[st + 1] += tv->step_off[st];
  tv->max_mF = 0;
  for (int k = 0; k < n; k++) tv->max_mF = std::max(tv->max_mF, tv->msgs[k].mF);
  // tile-walk form
  tv->tw.clear(); tv->stage_off.clear(); tv->step_stage.clear();
  bool fits32 = p->nslots_state < (int64_t)0xffffffffLL && p->nslots_resid < (int64_t)0xffffffffLL;
  if (n > 0 && tv->max_mF <= PGBP_TW_MAXM && fits32) {
    tv->tw.resize(n);
    for (int k = 0; k < n; k++) {
      const MsgDesc& m = tv->msgs[k];
      TwDesc d;
      memset(&d, 0, sizeof d);
      const int I = m.mF - m.s, S = m.s, SM = tri(m.mF), SS = tri(S);
      const int32_t* gat = p->tab.data() + m.gat;
      const int32_t* sca = p->tab.data() + m.sca;
      for (int q = 0; q < SM; q++) d.fJ[q] = (uint32_t)(m.fJ + gat[q]);
      for (int q = 0; q < m.mF; q++) d.fh[q] = (uint32_t)(m.fh + gat[SM + q]);
      for (int q = 0; q < SS; q++) d.tJ[q] = (uint32_t)(m.tJ + sca[q]);
      for (int q = 0; q < S; q++) d.th[q] = (uint32_t)(m.th + sca[SS + q]);
      d.fg = (uint32_t)m.fg; d.sg = (uint32_t)m.sg; d.tg = (uint32_t)m.tg;
      d.sJ = (uint32_t)m.sJ; d.sh = (uint32_t)m.sh; d.rJ = (uint32_t)m.rJ; d.rh = (uint32_t)m.rh;
      d.dmsg = (uint32_t)m.dmsg; d.ref = (uint32_t)m.ref; d.shape = (uint32_t)(I * 8 + S);
      tv->tw[k] = d;
    }
    tv->step_stage.assign(nsteps + 1, 0);
    tv->stage_off.push_back(0);
    for (int st = 0; st < nsteps; st++) {
      tv->step_stage[st] = (int32_t)tv->stage_off.size() - 1;
      for (int m0 = tv->step_off[st]; m0 < tv->step_off[st + 1]; m0 += PGBP_TW_STAGE)
        tv->stage_off.push_back(std::min(m0 + PGBP_TW_STAGE, tv->step_off[st + 1]));
    }
    tv->step_stage[nsteps] = (int32_t)tv->stage_off.size() - 1;
  }
  return 0;
}

extern "C" {

int32_t pgbp_abi_version(void) { return PGBP_ABI_VERSION; }

int32_t pgbp_last_error(char* buf, size_t buflen) {
  if (!buf || !buflen) return PGBP_EINVAL;
  const std::string& e = pgbp::last_error();
  size_t n = std::min(buflen - 1, e.size());
  memcpy(buf, e.data(), n);
  buf[n] = 0;
  return 0;
}

int32_t pgbp_plan_create(const pgbp_plan_desc* d, pgbp_plan** out) {
  if (!d || !out) PGBP_FAIL(PGBP_EINVAL, "null argument");
  if (d->nclusters <= 0 || d->nsepsets < 0 || d->ntraits <= 0) PGBP_FAIL(PGBP_EINVAL, "bad sizes");
  std::unique_ptr<pgbp_plan> p(new pgbp_plan);
  p->nclusters = d->nclusters;
  p->nsepsets = d->nsepsets;
  p->nbeliefs = d->nclusters + d->nsepsets;
  p->ntraits = d->ntraits;
  p->dim.assign(d->belief_dim, d->belief_dim + p->nbeliefs);
  p->jslot.resize(p->nbeliefs); p->hslot.resize(p->nbeliefs); p->gslot.resize(p->nbeliefs);
  int64_t slot = 0;
  for (int b = 0; b < p->nbeliefs; b++) {
    const int m = p->dim[b];
    if (m < 0 || m > PGBP_MAX_DIM) PGBP_FAIL(PGBP_EINVAL, "belief %d has dimension %d (max %d)", b, m, PGBP_MAX_DIM);
    p->max_dim = std::max(p->max_dim, m);
    p->jslot[b] = slot; slot += tri(m);
    p->hslot[b] = slot; slot += m;
    p->gslot[b] = slot; slot += 1;
    if (b == p->nclusters - 1) p->nslots_factor = slot;
  }
  p->nslots_state = slot;
  p->sep_a.resize(p->nsepsets); p->sep_b.resize(p->nsepsets);
  p->up_a.resize(p->nsepsets); p->up_b.resize(p->nsepsets);
  p->rjslot.resize(2 * (size_t)p->nsepsets); p->rhslot.resize(2 * (size_t)p->nsepsets);
  p->nbrs.resize(p->nclusters);
  int64_t rslot = 0;
  for (int j = 0; j < p->nsepsets; j++) {
    const int a = d->sepset_clusters[2 * j], b = d->sepset_clusters[2 * j + 1];
    if (a < 0 || a >= p->nclusters || b < 0 || b >= p->nclusters || a == b)
      PGBP_FAIL(PGBP_EINVAL, "sepset %d joins invalid clusters (%d,%d)", j, a, b);
    p->sep_a[j] = a; p->sep_b[j] = b;
    if (!p->sep_of.emplace(std::make_pair(std::min(a, b), std::max(a, b)), j).second)
      PGBP_FAIL(PGBP_EINVAL, "two sepsets join clusters (%d,%d)", a, b);
    const int s = p->dim[p->nclusters + j];
    for (int side = 0; side < 2; side++) {
      const int o0 = d->upind_off[2 * j + side], o1 = d->upind_off[2 * j + side + 1];
      const int c = side ? b : a;
      if (o1 - o0 != s) PGBP_FAIL(PGBP_EINVAL, "sepset %d: upind length %d != dimension %d", j, o1 - o0, s);
      std::vector<int32_t>& up = side ? p->up_b[j] : p->up_a[j];
      up.assign(d->upind + o0, d->upind + o1);
      for (int k = 0; k < s; k++) {
        if (up[k] < 0 || up[k] >= p->dim[c] || (k && up[k] <= up[k - 1]))
          PGBP_FAIL(PGBP_EINVAL, "sepset %d: upind into cluster %d not ascending / out of range", j, c);
      }
      p->rjslot[2 * j + side] = rslot; rslot += tri(s);
      p->rhslot[2 * j + side] = rslot; rslot += s;
    }
    p->nbrs[a].push_back({b, j});
    p->nbrs[b].push_back({a, j});
  }
  p->nslots_resid = rslot;
  for (auto& v : p->nbrs) std::sort(v.begin(), v.end());
  for (int j = 0; j < p->nsepsets; j++) {  // both directed messages of every sepset (see make_msg)
    std::vector<int32_t> g, sc;
    const int s = p->dim[p->nclusters + j];
    msg_tables(p.get(), p->dim[p->sep_a[j]], s, p->up_a[j], p->up_b[j], &g, &sc);
    p->intern_table(g); p->intern_table(sc);
    msg_tables(p.get(), p->dim[p->sep_b[j]], s, p->up_b[j], p->up_a[j], &g, &sc);
    p->intern_table(g); p->intern_table(sc);
  }
  // trees
  if (d->ntrees < 0) PGBP_FAIL(PGBP_EINVAL, "ntrees < 0");
  p->trees.resize(d->ntrees);
  for (int t = 0; t < d->ntrees; t++) {
    Tree& tr = p->trees[t];
    const int o0 = d->tree_off[t], o1 = d->tree_off[t + 1];
    const int n = o1 - o0;
    tr.parent.assign(d->tree_parent + o0, d->tree_parent + o1);
    tr.child.assign(d->tree_child + o0, d->tree_child + o1);
    tr.sepset.resize(n);
    for (int i = 0; i < n; i++) {
      const int a = tr.parent[i], b = tr.child[i];
      auto it = p->sep_of.find(std::make_pair(std::min(a, b), std::max(a, b)));
      if (it == p->sep_of.end()) PGBP_FAIL(PGBP_EINVAL, "tree %d edge %d: clusters (%d,%d) are not adjacent", t, i, a, b);
      tr.sepset[i] = it->second;
    }
    {
      std::set<int32_t> distinct(tr.sepset.begin(), tr.sepset.end());
      tr.covers_sepsets = n == p->nsepsets && (int)distinct.size() == n;
    }
    // postorder: i = n-1..0, child -> parent (src/calibration.jl:121-125)
    std::vector<int32_t> f(n), s(n), to(n);
    for (int r = 0; r < n; r++) {
      const int i = n - 1 - r;
      f[r] = tr.child[i]; s[r] = tr.sepset[i]; to[r] = tr.parent[i];
    }
    PGBP_TRY(build_traversal(p.get(), f, s, to, &tr.trav[0]));
    // preorder: i = 0..n-1, parent -> child (src/calibration.jl:147-151)
    for (int i = 0; i < n; i++) { f[i] = tr.parent[i]; s[i] = tr.sepset[i]; to[i] = tr.child[i]; }
    PGBP_TRY(build_traversal(p.get(), f, s, to, &tr.trav[1]));
    // walk list: reference order, postorder then preorder
    tr.walk.resize(2 * (size_t)n);
    tr.walkable = n > 0;
    for (int dir = 0; dir < 2; dir++)
      for (const MsgDesc& m : tr.trav[dir].msgs) {
        MsgDesc w = m;
        w.ref = m.ref + dir * n;
        tr.walk[w.ref] = w;
        if (w.wid < 0) tr.walkable = false;
      }
  }
  // node families
  if (d->families) {
    const pgbp_family_table* ft = d->families;
    FamilyTable& F = p->fam;
    F.nnodes = ft->nnodes; F.ntips = ft->ntips; F.root_fixed = ft->root_fixed;
    if (F.nnodes <= 0 || F.ntips < 0) PGBP_FAIL(PGBP_EINVAL, "bad family table sizes");
    F.node_cluster.assign(ft->node_cluster, ft->node_cluster + F.nnodes);
    F.mem_off.assign(ft->mem_off, ft->mem_off + F.nnodes + 1);
    const int nm = F.mem_off[F.nnodes];
    F.mem_pos.assign(ft->mem_pos, ft->mem_pos + nm);
    F.mem_length.assign(ft->mem_length, ft->mem_length + nm);
    F.mem_gamma.assign(ft->mem_gamma, ft->mem_gamma + nm);
    F.mem_color.assign(ft->mem_color, ft->mem_color + nm);
    F.node_datarow.assign(ft->node_datarow, ft->node_datarow + F.nnodes);
    const int pt = p->ntraits;
    F.scoped = ft->mem_tpos != nullptr || ft->tip_missing != nullptr;
    if (F.scoped) {
      if (ft->mem_tpos) F.mem_tpos.assign(ft->mem_tpos, ft->mem_tpos + (size_t)nm * pt);
      else {
        F.mem_tpos.assign((size_t)nm * pt, -1);
        for (int k = 0; k < nm; k++)
          for (int t = 0; t < pt; t++) if (F.mem_pos[k] >= 0) F.mem_tpos[(size_t)k * pt + t] = F.mem_pos[k] + t;
      }
      if (ft->tip_missing) F.tip_missing.assign(ft->tip_missing, ft->tip_missing + (size_t)F.ntips * pt);
      else F.tip_missing.assign((size_t)F.ntips * pt, 0);
    }
    std::vector<std::vector<int32_t>> c2n(p->nclusters);
    for (int v = 0; v < F.nnodes; v++) {
      const int c = F.node_cluster[v];
      if (c < 0 || c >= p->nclusters) PGBP_FAIL(PGBP_EINVAL, "node %d assigned to invalid cluster %d", v, c);
      const int o0 = F.mem_off[v], o1 = F.mem_off[v + 1];
      if (o1 - o0 < 1 || o1 - o0 > PGBP_MAX_FAMILY) PGBP_FAIL(PGBP_EINVAL, "node %d: family size %d unsupported", v, o1 - o0);
      for (int k = o0; k < o1; k++) {
        if (F.scoped) {
          for (int t = 0; t < pt; t++)
            if (F.mem_tpos[(size_t)k * pt + t] >= p->dim[c]) PGBP_FAIL(PGBP_EINVAL, "node %d: member scope exceeds cluster %d", v, c);
        } else if (F.mem_pos[k] >= 0 && F.mem_pos[k] + pt > p->dim[c])
          PGBP_FAIL(PGBP_EINVAL, "node %d: member scope exceeds cluster %d", v, c);
        if (k > o0) {
          if (!(F.mem_length[k] >= 0)) PGBP_FAIL(PGBP_EINVAL, "node %d: negative edge length", v);
          if (F.mem_color[k] < 0) PGBP_FAIL(PGBP_EINVAL, "node %d: negative colour", v);
          F.ncolors_min = std::max(F.ncolors_min, F.mem_color[k] + 1);
        }
      }
      if (F.node_datarow[v] >= F.ntips) PGBP_FAIL(PGBP_EINVAL, "node %d: data row out of range", v);
      if (F.scoped && (o1 - o0) * pt > PGBP_SCOPED_MAXN)
        PGBP_FAIL(PGBP_EINVAL, "node %d: family of %d members x %d traits exceeds the %d variables of the scoped "
                  "factor-assignment path", v, o1 - o0, pt, PGBP_SCOPED_MAXN);
      c2n[c].push_back(v);
    }
    F.clu_off.assign(1, 0);
    F.clu_flag.assign(p->nclusters, 0);
    F.first_J.assign(F.nnodes, 0);
    F.first_h.assign(F.nnodes, 0);
    for (int c = 0; c < p->nclusters; c++) {
      auto& v = c2n[c];
      F.clu_node.insert(F.clu_node.end(), v.begin(), v.end());
      F.clu_off.push_back((int32_t)F.clu_node.size());
      std::set<std::pair<int, int>> seenJ;  // (pos_a, pos_b) blocks already written
      std::set<int> seenh;
      for (int node : v) {
        if (F.scoped) break;  // the scoped body zero-fills and accumulates
        const int o0 = F.mem_off[node], nm = F.mem_off[node + 1] - o0;
        for (int a = 0; a < nm; a++) {
          const int pa = F.mem_pos[o0 + a];
          if (pa < 0) continue;
          if (seenh.insert(pa).second) F.first_h[node] |= (uint8_t)(1u << a);
          for (int bq = a; bq < nm; bq++) {
            const int pb = F.mem_pos[o0 + bq];
            if (pb < 0) continue;
            if (pb < pa) PGBP_FAIL(PGBP_EINVAL, "node %d: family members are not in cluster order", node);
            if (seenJ.insert({pa, pb}).second) F.first_J[node] |= (uint64_t)1 << (a * 8 + bq);
          }
        }
      }
      const int nn = p->dim[c] / pt;  // full trait scopes: in-scope nodes of the cluster
      if ((int)seenh.size() != nn || (int)seenJ.size() != nn * (nn + 1) / 2) F.clu_flag[c] |= 1;
    }
    p->has_families = true;
  }
  *out = p.release();
  return 0;
}

int32_t pgbp_plan_destroy(pgbp_plan* plan) {
  delete plan;
  return 0;
}

int32_t pgbp_plan_get_levels(const pgbp_plan* plan, int32_t tree, int32_t direction, int32_t* nmsg,
                             int32_t* nsteps, int32_t* msg_ref, int32_t* msg_step, int32_t* msg_from,
                             int32_t* msg_sepset, int32_t* msg_to) {
  if (!plan || tree < 0 || tree >= (int)plan->trees.size() || direction < 0 || direction > 1)
    PGBP_FAIL(PGBP_EINVAL, "bad tree / direction");
  const Tree& tr = plan->trees[tree];
  const Traversal& tv = tr.trav[direction];
  const int n = (int)tv.msgs.size();
  if (nmsg) *nmsg = n;
  if (nsteps) *nsteps = tv.nsteps;
  for (int k = 0; k < n; k++) {
    const int r = tv.msgs[k].ref;
    const int i = direction == 0 ? n - 1 - r : r;
    if (msg_ref) msg_ref[k] = r;
    if (msg_step) msg_step[k] = tv.step_of_msg[k];
    if (msg_from) msg_from[k] = direction == 0 ? tr.child[i] : tr.parent[i];
    if (msg_to) msg_to[k] = direction == 0 ? tr.parent[i] : tr.child[i];
    if (msg_sepset) msg_sepset[k] = plan->nclusters + tr.sepset[i];
  }
  return 0;
}

int32_t pgbp_plan_traversal_cost(const pgbp_plan* plan, int32_t tree, int32_t direction,
                                 int32_t track_residuals, double* bytes, double* flops) {
  if (!plan || tree < 0 || tree >= (int)plan->trees.size() || direction < 0 || direction > 1)
    PGBP_FAIL(PGBP_EINVAL, "bad tree / direction");
  const Traversal& tv = plan->trees[tree].trav[direction];
  if (bytes) *bytes = track_residuals ? tv.bytes_resid : tv.bytes_noresid;
  if (flops) *flops = tv.flops;
  return 0;
}

int32_t pgbp_belief_slot(const pgbp_plan* plan, int32_t belief, int64_t* jslot, int64_t* hslot, int64_t* gslot) {
  if (!plan || belief < 0 || belief >= plan->nbeliefs) PGBP_FAIL(PGBP_EINVAL, "bad belief index");
  if (jslot) *jslot = plan->jslot[belief];
  if (hslot) *hslot = plan->hslot[belief];
  if (gslot) *gslot = plan->gslot[belief];
  return 0;
}

}  // extern "C"
