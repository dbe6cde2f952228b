/* hbt_oracle.c - plain-C CPU restatement of HBT+'s unbinding path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load libhbtoracle.so; the product
 * (hbtplus_b200/) never does and has no CPU fallback.
 *
 * PARITY PIN: this restatement is checked against the UNMODIFIED reference sources
 * compiled as oracle/_ref/libhbtref_v32.so (tests/test_oracle.py, run in the build
 * container) and against the committed fixtures in tests/golden/ that were generated
 * from that library (tests/golden/make_golden.py).  The reference itself ships no
 * golden vectors for this path.
 *
 * V32 ABI: HBTInt=int32, HBTReal=float, DM_ONLY.  Every function cites the reference
 * lines it follows (paths relative to /root/reference).  Arithmetic widths (float vs
 * double) are those of the reference expressions; compile with -ffp-contract=off.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hbt_unbind.h"

typedef int32_t HBTInt;
typedef float HBTReal;

typedef struct
{ /* Particle_t, src/snapshot.h:51-72 (DM_ONLY) */
  HBTInt Id;
  HBTReal x[3], v[3], m;
  HBTReal u; /* InternalEnergy (HAS_THERMAL_ENERGY builds, src/snapshot.h:58-60) */
} Particle;

typedef struct
{ /* ParticleEnergy_t, src/subhalo_unbind.cpp:12-16 */
  HBTInt pid;
  float E;
} PE;

typedef struct
{ /* Parameter_t subset + PhysicalConst::G, in the reference's storage types */
  int MinNumPartOfSub, Periodic, RefineMostbound;
  HBTInt MaxSampleSize;
  HBTReal BoundMassPrecision, SourceSubRelaxFactor, BoxSize, BoxHalf, SofteningHalo, OpenAngleSquare, TreeNodeResolution,
      TreeNodeResolutionHalf, G;
  HBTReal ScaleFactor, Hz;
  int SnapshotIndex;
  int64_t ShuffleSeed;
  int NoStripping, ThermalEnergy; /* the -DNO_STRIPPING / -DUNBIND_WITH_THERMAL_ENERGY builds (batch flags) */
} Config;

static void config_from(Config *c, const hbtu_params *p, const hbtu_epoch *e)
{
  c->MinNumPartOfSub = p->min_num_part_of_sub;
  c->Periodic = p->periodic_boundary_on;
  c->RefineMostbound = p->refine_mostbound_particle;
  c->MaxSampleSize = (HBTInt)p->max_sample_size;
  c->BoundMassPrecision = (HBTReal)p->bound_mass_precision;
  c->SourceSubRelaxFactor = (HBTReal)p->source_sub_relax_factor;
  c->BoxSize = (HBTReal)p->box_size;
  c->BoxHalf = (HBTReal)p->box_half;
  c->SofteningHalo = (HBTReal)p->softening_halo;
  c->OpenAngleSquare = (HBTReal)p->tree_node_open_angle_square;
  c->TreeNodeResolution = (HBTReal)p->tree_node_resolution;
  c->TreeNodeResolutionHalf = (HBTReal)p->tree_node_resolution_half;
  c->G = (HBTReal)p->G;
  c->ScaleFactor = (HBTReal)e->scale_factor;
  c->Hz = (HBTReal)e->hz;
  c->SnapshotIndex = e->snapshot_index;
  c->ShuffleSeed = p->shuffle_seed;
  c->NoStripping = c->ThermalEnergy = 0;
}

/* NEAREST(), src/config_parser.h:142 - one instance per operand width */
static inline float nearest_f(const Config *c, float x) { return x > c->BoxHalf ? x - c->BoxSize : (x < -c->BoxHalf ? x + c->BoxSize : x); }
static inline double nearest_d(const Config *c, double x) { return x > c->BoxHalf ? x - c->BoxSize : (x < -c->BoxHalf ? x + c->BoxSize : x); }

/* ---------------------------------------------------------------------------------------------
 * Source view: EnergySnapshot_t (src/subhalo_unbind.cpp:68-107): i -> Particles[Elist[i].pid],
 * mass scaled by MassFactor.  With Elist==NULL it is a direct view (i -> P[i]).
 * ------------------------------------------------------------------------------------------- */
typedef struct
{
  const Particle *P;
  const PE *Elist;
  HBTReal MassFactor;
} View;
static inline const Particle *view_p(const View *v, HBTInt i) { return v->Elist ? &v->P[v->Elist[i].pid] : &v->P[i]; }
static inline HBTReal view_mass(const View *v, HBTInt i) { return view_p(v, i)->m * v->MassFactor; }

/* ---------------------------------------------------------------------------------------------
 * Octree: OctTree_t<GravityTreeCell_t>, src/oct_tree.h:26-70, src/oct_tree.tpp:9-157.
 * The reference's union TreeCell_t {sons[8] | way{s[3],len,mass,sibling,nextnode}} is kept as
 * two parallel arrays (sons are backed up before way is written, src/gravity_tree.cpp:55-57).
 * ------------------------------------------------------------------------------------------- */
typedef struct
{
  HBTReal s[3], len, mass;
  HBTInt sibling, nextnode;
} Way;
typedef struct
{
  HBTInt np, ncell, cap;
  HBTInt (*sons)[8];
  Way *way;
  HBTInt *nextnode_from_particle;
  View view;
  const Config *cfg;
} Tree;

static void tree_free(Tree *t)
{
  free(t->sons);
  free(t->way);
  free(t->nextnode_from_particle);
  memset(t, 0, sizeof(*t));
}
static HBTInt tree_append_cell(Tree *t)
{ /* AppendCell, src/oct_tree.tpp:9-14 */
  if (t->ncell == t->cap)
  {
    t->cap = t->cap ? t->cap * 2 : 16;
    t->sons = realloc(t->sons, sizeof(*t->sons) * t->cap);
    t->way = realloc(t->way, sizeof(*t->way) * t->cap);
  }
  for (int j = 0; j < 8; j++) t->sons[t->ncell][j] = -1;
  return t->ncell++;
}
#define SONS(t, nodeid) ((t)->sons[(nodeid) - (t)->np])
#define WAY(t, nodeid) ((t)->way[(nodeid) - (t)->np])

static void update_internal_nodes(Tree *t, HBTInt no, HBTInt sib, double len);

static void process_node(Tree *t, HBTInt nodeid, HBTInt nextid, double *mass, double CoM[3], double len)
{ /* ProcessNode, src/gravity_tree.cpp:25-46 */
  if (nodeid < t->np)
  {
    double thismass = view_mass(&t->view, nodeid);
    const Particle *p = view_p(&t->view, nodeid);
    *mass += thismass;
    for (int j = 0; j < 3; j++) CoM[j] += p->x[j] * thismass;
    t->nextnode_from_particle[nodeid] = nextid;
  }
  else
  {
    if (len >= t->cfg->TreeNodeResolution)
      update_internal_nodes(t, nodeid, nextid, len / 2.);
    else
      update_internal_nodes(t, nodeid, nextid, len);
    double thismass = WAY(t, nodeid).mass;
    *mass += thismass;
    for (int j = 0; j < 3; j++) CoM[j] += WAY(t, nodeid).s[j] * thismass;
  }
}
static void update_internal_nodes(Tree *t, HBTInt no, HBTInt sib, double len)
{ /* UpdateInternalNodes + FillNodeCenter, src/gravity_tree.cpp:18-23,48-77 */
  HBTInt p, pp, sons[8];
  int j, jj, i;
  double mass = 0., CoM[3] = {0., 0., 0.};
  for (j = 0; j < 8; j++) sons[j] = SONS(t, no)[j];
  WAY(t, no).len = len;
  WAY(t, no).sibling = sib;
  for (i = 0; sons[i] < 0; i++)
    ;
  jj = i;
  pp = sons[jj];
  WAY(t, no).nextnode = pp;
  for (i++; i < 8; i++)
    if (sons[i] >= 0)
    {
      j = jj;
      p = pp;
      jj = i;
      pp = sons[jj];
      process_node(t, p, pp, &mass, CoM, len);
    }
  (void)j;
  process_node(t, pp, sib, &mass, CoM, len);
  WAY(t, no).mass = mass;
  WAY(t, no).s[0] = CoM[0] / mass;
  WAY(t, no).s[1] = CoM[1] / mass;
  WAY(t, no).s[2] = CoM[2] / mass;
}

static HBTInt tree_build(Tree *t, const Config *cfg, View view, HBTInt num_part)
{ /* OctTree_t::Build, src/oct_tree.tpp:17-144 */
  HBTInt sub, subid, i, nodeid;
  int j;
  double center[3], lenhalf, xmin[3], xmax[3], Center[3], Len, Lenhalf;
  t->cfg = cfg;
  t->view = view;
  t->np = num_part;
  t->ncell = 0;
  t->nextnode_from_particle = realloc(t->nextnode_from_particle, sizeof(HBTInt) * (num_part > 0 ? num_part : 1));
  for (j = 0; j < 3; j++) xmin[j] = xmax[j] = view_p(&view, 0)->x[j];
  for (i = 1; i < num_part; i++)
    for (j = 0; j < 3; j++)
    {
      HBTReal x = view_p(&view, i)->x[j];
      if (x > xmax[j])
        xmax[j] = x;
      else if (x < xmin[j])
        xmin[j] = x;
    }
  for (j = 1, Len = xmax[0] - xmin[0]; j < 3; j++)
    if ((xmax[j] - xmin[j]) > Len) Len = xmax[j] - xmin[j];
  for (j = 0; j < 3; j++) Center[j] = 0.5 * (xmax[j] + xmin[j]);
  Lenhalf = 0.5 * Len;
  tree_append_cell(t); /* root = node id np */
  for (i = 0; i < num_part; i++)
  {
    const HBTReal *xi = view_p(&view, i)->x;
    nodeid = t->np;
    lenhalf = Lenhalf;
    for (j = 0; j < 3; j++) center[j] = Center[j];
    while (1)
    {
      lenhalf *= 0.5;
      sub = 0;
      for (j = 0; j < 3; j++)
        if (xi[j] > center[j])
        {
          center[j] += lenhalf;
          sub += 1 << j;
        }
        else
          center[j] -= lenhalf;
      subid = SONS(t, nodeid)[sub];
      if (subid < 0)
      {
        SONS(t, nodeid)[sub] = i;
        break;
      }
      else if (subid < t->np)
      {
        HBTInt newnodeid = t->np + t->ncell;
        tree_append_cell(t);
        SONS(t, nodeid)[sub] = newnodeid;
        nodeid = newnodeid;
        if (lenhalf < cfg->TreeNodeResolutionHalf)
        { /* co-located particles: random octant, src/oct_tree.tpp:112-123 */
          sub = (HBTInt)(8.0 * drand48());
          if (sub >= 8) sub = 7;
        }
        else
        {
          const HBTReal *xs = view_p(&view, subid)->x;
          sub = 0;
          for (j = 0; j < 3; j++)
            if (xs[j] > center[j]) sub += 1 << j;
        }
        SONS(t, nodeid)[sub] = subid;
      }
      else
        nodeid = subid;
    }
  }
  update_internal_nodes(t, t->np, -1, Len);
  return t->ncell;
}

/* EvaluatePotential, src/gravity_tree.cpp:79-164.  counts[0]+=accepted sources (particles and
 * accepted nodes), counts[1]+=opened nodes, when counts!=NULL. */
static double tree_potential(const Tree *t, const HBTReal targetPos[3], HBTReal targetMass, int64_t *counts)
{
  const Config *c = t->cfg;
  HBTInt no;
  double r2, dx, dy, dz, mass, r, u, h, h_inv, wp, pot;
  double pos_x = targetPos[0], pos_y = targetPos[1], pos_z = targetPos[2];
  int64_t nacc = 0, nopen = 0;
  h = 2.8 * c->SofteningHalo;
  h_inv = 1.0 / h;
  pot = targetMass / c->SofteningHalo;
  no = t->np;
  while (no >= 0)
  {
    if (no < t->np)
    {
      const Particle *p = view_p(&t->view, no);
      dx = p->x[0] - pos_x;
      dy = p->x[1] - pos_y;
      dz = p->x[2] - pos_z;
      if (c->Periodic)
      {
        dx = nearest_d(c, dx);
        dy = nearest_d(c, dy);
        dz = nearest_d(c, dz);
      }
      mass = view_mass(&t->view, no);
      no = t->nextnode_from_particle[no];
      r2 = dx * dx + dy * dy + dz * dz;
    }
    else
    {
      const Way *nop = &WAY(t, no);
      dx = nop->s[0] - pos_x;
      dy = nop->s[1] - pos_y;
      dz = nop->s[2] - pos_z;
      if (c->Periodic)
      {
        dx = nearest_d(c, dx);
        dy = nearest_d(c, dy);
        dz = nearest_d(c, dz);
      }
      mass = nop->mass;
      r2 = dx * dx + dy * dy + dz * dz;
      if ((nop->len * nop->len) > (r2 * c->OpenAngleSquare)) /* float*float vs double*float */
      {
        no = nop->nextnode;
        nopen++;
        continue;
      }
      no = nop->sibling;
    }
    nacc++;
    r = sqrt(r2);
    if (r >= h)
      pot -= mass / r;
    else
    {
      u = r * h_inv;
      if (u < 0.5)
        wp = -2.8 + u * u * (5.333333333333 + u * u * (6.4 * u - 9.6));
      else
        wp = -3.2 + 0.066666666667 / u + u * u * (10.666666666667 + u * (-16.0 + u * (9.6 - 2.133333333333 * u)));
      pot += mass * h_inv * wp;
    }
  }
  if (counts)
  {
    counts[0] += nacc;
    counts[1] += nopen;
  }
  return pot * c->G / c->ScaleFactor;
}

static void relative_velocity(const Config *c, const HBTReal tp[3], const HBTReal tv[3], const HBTReal rp[3], const HBTReal rv[3],
                              HBTReal dv[3])
{ /* Snapshot_t::RelativeVelocity, src/snapshot.h:100-111 (all HBTReal arithmetic) */
  for (int j = 0; j < 3; j++)
  {
    HBTReal dx = tp[j] - rp[j];
    if (c->Periodic) dx = nearest_f(c, dx);
    dv[j] = tv[j] - rv[j];
    dv[j] += c->Hz * c->ScaleFactor * dx;
  }
}
static double binding_energy(const Tree *t, const HBTReal tp[3], const HBTReal tv[3], const HBTReal rp[3], const HBTReal rv[3],
                             HBTReal targetMass, int64_t *counts)
{ /* GravityTree_t::BindingEnergy, src/gravity_tree.cpp:166-175 */
  double pot = tree_potential(t, tp, targetMass, counts);
  HBTReal dv[3];
  relative_velocity(t->cfg, tp, tv, rp, rv, dv);
  return (dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]) * 0.5 + pot;
}

/* ---------------------------------------------------------------------------------------------
 * Subhalo_t subset, src/subhalo.h:24-146
 * ------------------------------------------------------------------------------------------- */
typedef struct
{
  Particle *P;
  int64_t n, cap;
  HBTInt Nbound;
  float Mbound;
  HBTReal AvgPos[3], AvgVel[3], MbPos[3], MbVel[3];
  int Death, Sink;
  HBTInt SinkTrackId;
  float SpecPot, SpecKin, SpecAM[3];
  float *Energies;
  int64_t nE;
  int iterations;
  int64_t index; /* batch-local subhalo index */
  const int32_t *nest;
  int64_t nnest;
} Sub;

/* Sampling permutation.  mode 0: libstdc++ random_shuffle on libc rand(), bit-identical to the reference when both run
 * single-threaded from the same srand() state (pins this restatement to the reference).  mode 1: the counter-based
 * permutation the CUDA path uses (sort positions by shuffle_key; same formula as hbtplus_b200/csrc/tree_core.cuh),
 * because the reference's own stream depends on libc state and OpenMP scheduling and cannot be shared with a GPU. */
static int g_shuffle_mode = 0;
static uint64_t shuffle_key(uint64_t seed, uint64_t sub, uint64_t j)
{
  uint64_t z = seed ^ (0x9E3779B97F4A7C15ULL * (sub + 1)) ^ (j * 0xBF58476D1CE4E5B9ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return z >> 24; /* 40 bits: shares a 64-bit radix key with the segment index on the device */
}
typedef struct
{
  uint64_t key;
  int64_t j;
} ShuffleRec;
static int comp_shuffle(const void *a, const void *b)
{
  const ShuffleRec *x = a, *y = b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->j < y->j ? -1 : (x->j > y->j);
}

static int64_t g_interactions; /* accepted pair interactions of the last call (roofline numerator) */
static int64_t g_opened;

static int comp_energy(const void *a, const void *b)
{ /* CompEnergy, src/subhalo_unbind.cpp:17-20 */
  float ea = ((const PE *)a)->E, eb = ((const PE *)b)->E;
  return (ea < eb) ? -1 : (ea > eb);
}

static HBTInt partition_binding_energy(PE *Elist, size_t len)
{ /* PartitionBindingEnergy, src/subhalo_unbind.cpp:21-58 (hole-based Hoare partition) */
  if (len == 0) return 0;
  if (len == 1) return Elist[0].E < 0;
  PE Etmp = Elist[0];
  PE *f = Elist, *b = Elist + len;
  while (1)
  {
    while (1)
    {
      b--;
      if (b == f)
      {
        *f = Etmp;
        if (Etmp.E < 0) b++;
        return (HBTInt)(b - Elist);
      }
      if (b->E < 0) break;
    }
    *f = *b;
    while (1)
    {
      f++;
      if (f == b)
      {
        *b = Etmp;
        if (Etmp.E < 0) b++;
        return (HBTInt)(b - Elist);
      }
      if (f->E > 0) break;
    }
    *b = *f;
  }
}

static double average_velocity(const View *v, HBTReal CoV[3], HBTInt n)
{ /* EnergySnapshot_t::AverageVelocity, src/subhalo_unbind.cpp:108-138 */
  if (0 == n) return 0.;
  if (1 == n)
  {
    for (int j = 0; j < 3; j++) CoV[j] = view_p(v, 0)->v[j];
    return view_mass(v, 0);
  }
  double svx = 0, svy = 0, svz = 0, msum = 0;
  for (HBTInt i = 0; i < n; i++)
  {
    HBTReal m = view_mass(v, i);
    const HBTReal *vel = view_p(v, i)->v;
    msum += m;
    svx += vel[0] * m; /* float product, double accumulate */
    svy += vel[1] * m;
    svz += vel[2] * m;
  }
  CoV[0] = svx / msum;
  CoV[1] = svy / msum;
  CoV[2] = svz / msum;
  return msum;
}
static double average_position(const Config *c, const View *v, HBTReal CoM[3], HBTInt n)
{ /* EnergySnapshot_t::AveragePosition, src/subhalo_unbind.cpp:139-188 */
  if (0 == n) return 0.;
  if (1 == n)
  {
    for (int j = 0; j < 3; j++) CoM[j] = view_p(v, 0)->x[j];
    return view_mass(v, 0);
  }
  double sx = 0, sy = 0, sz = 0, origin[3] = {0, 0, 0}, msum = 0;
  if (c->Periodic)
    for (int j = 0; j < 3; j++) origin[j] = view_p(v, 0)->x[j];
  for (HBTInt i = 0; i < n; i++)
  {
    HBTReal m = view_mass(v, i);
    const HBTReal *x = view_p(v, i)->x;
    msum += m;
    if (c->Periodic)
    {
      sx += nearest_d(c, x[0] - origin[0]) * m;
      sy += nearest_d(c, x[1] - origin[1]) * m;
      sz += nearest_d(c, x[2] - origin[2]) * m;
    }
    else
    {
      sx += x[0] * m;
      sy += x[1] * m;
      sz += x[2] * m;
    }
  }
  sx /= msum;
  sy /= msum;
  sz /= msum;
  if (c->Periodic)
  {
    sx += origin[0];
    sy += origin[1];
    sz += origin[2];
  }
  CoM[0] = sx;
  CoM[1] = sy;
  CoM[2] = sz;
  return msum;
}
static void average_kinematics(const Config *c, const View *v, float *SpecPot, float *SpecKin, float AM[3], HBTInt n,
                               const HBTReal refPos[3], const HBTReal refVel[3])
{ /* EnergySnapshot_t::AverageKinematics, src/subhalo_unbind.cpp:189-232 */
  if (n <= 1)
  {
    *SpecPot = 0.;
    *SpecKin = 0.;
    AM[0] = AM[1] = AM[2] = 0.;
    return;
  }
  double E = 0., K = 0., AMx = 0., AMy = 0., AMz = 0., M = 0.;
  for (HBTInt i = 0; i < n; i++)
  {
    HBTReal m = view_mass(v, i);
    E += v->Elist[i].E * m;
    const Particle *p = view_p(v, i);
    double dx[3], dv[3];
    for (int j = 0; j < 3; j++)
    {
      dx[j] = p->x[j] - refPos[j];
      if (c->Periodic) dx[j] = nearest_d(c, dx[j]);
      dx[j] *= c->ScaleFactor;
      dv[j] = p->v[j] - refVel[j] + c->Hz * dx[j];
      K += dv[j] * dv[j] * m;
    }
    AMx += (dx[1] * dv[2] - dx[2] * dv[1]) * m;
    AMy += (dx[2] * dv[0] - dx[0] * dv[2]) * m;
    AMz += (dx[0] * dv[1] - dx[1] * dv[0]) * m;
    M += m;
  }
  E /= M;
  K *= 0.5 / M;
  *SpecPot = E - K;
  *SpecKin = K;
  AM[0] = AMx / M;
  AM[1] = AMy / M;
  AM[2] = AMz / M;
}

static void refine_binding_energy_order(const Config *c, View *ESnap, PE *Elist, HBTInt Size, Tree *tree, const HBTReal RefPos[3],
                                        const HBTReal RefVel[3])
{ /* RefineBindingEnergyOrder, src/subhalo_unbind.cpp:234-262 */
  tree_build(tree, c, *ESnap, Size);
  PE *Einner = malloc(sizeof(PE) * Size);
  int64_t cnt[2] = {0, 0};
  for (HBTInt i = 0; i < Size; i++)
  {
    const Particle *p = &ESnap->P[Elist[i].pid];
    Einner[i].pid = i;
    Einner[i].E = binding_energy(tree, p->x, p->v, RefPos, RefVel, p->m, cnt);
  }
  g_interactions += cnt[0];
  g_opened += cnt[1];
  qsort(Einner, Size, sizeof(PE), comp_energy);
  for (HBTInt i = 0; i < Size; i++) Einner[i] = Elist[Einner[i].pid];
  for (HBTInt i = 0; i < Size; i++) Elist[i] = Einner[i];
  free(Einner);
}

static void count_particles(Sub *s)
{ /* Subhalo_t::CountParticles (DM_ONLY), src/subhalo.cpp:479-487 */
  s->Mbound = 0.;
  for (HBTInt i = 0; i < s->Nbound; i++) s->Mbound += s->P[i].m;
}

static void unbind(const Config *c, Sub *s)
{ /* Subhalo_t::Unbind, src/subhalo_unbind.cpp:263-431 */
  HBTInt MaxSampleSize = c->MaxSampleSize;
  int RefineMostboundParticle = (MaxSampleSize > 0 && c->RefineMostbound);
  HBTReal BoundMassPrecision = c->BoundMassPrecision;
  s->iterations = 0;
  if (s->n < c->MinNumPartOfSub)
    if (s->Death == -1) s->Death = c->SnapshotIndex;
  if (s->n == 0)
  {
    s->Nbound = 0;
    count_particles(s);
    s->nE = 0;
    return;
  }
  if (s->n == 1)
  {
    s->Nbound = 1;
    count_particles(s);
    s->nE = 1;
    s->Energies = realloc(s->Energies, sizeof(float));
    s->Energies[0] = 0.;
    return;
  }
  HBTReal OldRefPos[3] = {0, 0, 0}, OldRefVel[3] = {0, 0, 0};
  HBTReal *RefPos = s->AvgPos, *RefVel = s->AvgVel;
  Particle OldMostbound = s->P[0];
  Tree tree;
  memset(&tree, 0, sizeof(tree));
  s->Nbound = (HBTInt)s->n;
  if (MaxSampleSize > 0 && s->Nbound > MaxSampleSize)
  { /* std::random_shuffle (libstdc++ stl_algo.h: j = rand() % (i+1)), src/subhalo_unbind.cpp:302 */
    if (g_shuffle_mode == 0)
    {
      for (int64_t i = 1; i < s->n; i++)
      {
        int64_t j = rand() % (i + 1);
        if (i != j)
        {
          Particle tmp = s->P[i];
          s->P[i] = s->P[j];
          s->P[j] = tmp;
        }
      }
    }
    else
    {
      ShuffleRec *rec = malloc(sizeof(ShuffleRec) * s->n);
      Particle *np = malloc(sizeof(Particle) * s->n);
      for (int64_t j = 0; j < s->n; j++)
      {
        rec[j].key = shuffle_key((uint64_t)c->ShuffleSeed, (uint64_t)s->index, (uint64_t)j);
        rec[j].j = j;
      }
      qsort(rec, s->n, sizeof(ShuffleRec), comp_shuffle);
      for (int64_t j = 0; j < s->n; j++) np[j] = s->P[rec[j].j];
      free(s->P);
      free(rec);
      s->P = np;
      s->cap = s->n;
    }
  }
  HBTInt Nlast = 0;
  PE *Elist = malloc(sizeof(PE) * s->n);
  for (HBTInt i = 0; i < s->Nbound; i++)
  {
    Elist[i].pid = i;
    Elist[i].E = 0;
  }
  View ESnap = {s->P, Elist, 1.f};
  int CorrectionLoop = 0;
  while (1)
  {
    s->iterations++;
    if (CorrectionLoop)
    { /* src/subhalo_unbind.cpp:312-330 */
      HBTReal RefVelDiff[3];
      relative_velocity(c, OldRefPos, OldRefVel, RefPos, RefVel, RefVelDiff);
      HBTReal dK = 0.5 * (RefVelDiff[0] * RefVelDiff[0] + RefVelDiff[1] * RefVelDiff[1] + RefVelDiff[2] * RefVelDiff[2]);
      View ESnapCorrection = {s->P, &Elist[s->Nbound], 1.f};
      tree_build(&tree, c, ESnapCorrection, Nlast - s->Nbound);
      int64_t cnt0 = 0, cnt1 = 0;
#pragma omp parallel for if (Nlast > 100) reduction(+ : cnt0, cnt1) schedule(dynamic, 64)
      for (HBTInt i = 0; i < s->Nbound; i++)
      {
        const Particle *p = &s->P[Elist[i].pid];
        HBTReal OldVel[3];
        int64_t cnt[2] = {0, 0};
        relative_velocity(c, p->x, p->v, OldRefPos, OldRefVel, OldVel);
        Elist[i].E += (OldVel[0] * RefVelDiff[0] + OldVel[1] * RefVelDiff[1] + OldVel[2] * RefVelDiff[2]) + dK -
                      tree_potential(&tree, p->x, 0, cnt);
        cnt0 += cnt[0];
        cnt1 += cnt[1];
      }
      g_interactions += cnt0;
      g_opened += cnt1;
      Nlast = s->Nbound;
    }
    else
    { /* src/subhalo_unbind.cpp:331-356 */
      Nlast = s->Nbound;
      HBTInt np_tree = Nlast;
      if (MaxSampleSize > 0 && Nlast > MaxSampleSize)
      {
        np_tree = MaxSampleSize;
        ESnap.MassFactor = (HBTReal)Nlast / MaxSampleSize;
      }
      tree_build(&tree, c, ESnap, np_tree);
      int64_t cnt0 = 0, cnt1 = 0;
#pragma omp parallel for if (Nlast > 100) reduction(+ : cnt0, cnt1) schedule(dynamic, 64)
      for (HBTInt i = 0; i < Nlast; i++)
      {
        const Particle *p = &s->P[Elist[i].pid];
        int64_t cnt[2] = {0, 0};
        HBTReal mass = (i < np_tree) ? view_mass(&ESnap, i) : 0.f;
        Elist[i].E = binding_energy(&tree, p->x, p->v, RefPos, RefVel, mass, cnt);
        if (c->ThermalEnergy) Elist[i].E += p->u; /* UNBIND_WITH_THERMAL_ENERGY, src/subhalo_unbind.cpp:351-353 */
        cnt0 += cnt[0];
        cnt1 += cnt[1];
      }
      g_interactions += cnt0;
      g_opened += cnt1;
      ESnap.MassFactor = 1.f;
    }
    s->Nbound = partition_binding_energy(Elist, Nlast);
    if (c->NoStripping) s->Nbound = Nlast; /* NO_STRIPPING, src/subhalo_unbind.cpp:358-360 */
    if (s->Nbound < c->MinNumPartOfSub)
    { /* disruption, src/subhalo_unbind.cpp:361-379 */
      s->Nbound = 1;
      Nlast = 1;
      if (s->Death == -1) s->Death = c->SnapshotIndex;
      for (int64_t i = 0; i < s->n; i++)
        if (s->P[i].Id == OldMostbound.Id)
        {
          Particle tmp = s->P[i];
          s->P[i] = s->P[0];
          s->P[0] = tmp;
          break;
        }
      memcpy(s->AvgPos, s->MbPos, sizeof(s->AvgPos));
      memcpy(s->AvgVel, s->MbVel, sizeof(s->AvgVel));
      s->Mbound = s->P[0].m;
      break;
    }
    else
    {
      qsort(Elist + s->Nbound, Nlast - s->Nbound, sizeof(PE), comp_energy);
      HBTInt Ndiff = Nlast - s->Nbound;
      if (Ndiff < s->Nbound)
        if (MaxSampleSize <= 0 || Ndiff < MaxSampleSize)
        {
          CorrectionLoop = 1;
          memcpy(OldRefPos, RefPos, sizeof(OldRefPos));
          memcpy(OldRefVel, RefVel, sizeof(OldRefVel));
        }
      s->Mbound = average_velocity(&ESnap, s->AvgVel, s->Nbound);
      average_position(c, &ESnap, s->AvgPos, s->Nbound);
      if (s->Nbound >= Nlast * BoundMassPrecision) /* int*float -> float compare */
      {
        if (s->Death != -1) s->Death = -1;
        if (s->SinkTrackId != -1)
        {
          s->Sink = -1;
          s->SinkTrackId = -1;
        }
        qsort(Elist, s->Nbound, sizeof(PE), comp_energy);
        if (RefineMostboundParticle && s->Nbound > MaxSampleSize)
          refine_binding_energy_order(c, &ESnap, Elist, MaxSampleSize, &tree, RefPos, RefVel);
        Particle *p = malloc(sizeof(Particle) * s->n);
        for (int64_t i = 0; i < s->n; i++)
        {
          p[i] = s->P[Elist[i].pid];
          Elist[i].pid = (HBTInt)i;
        }
        free(s->P);
        s->P = p;
        s->cap = s->n;
        ESnap.P = p;
        memcpy(s->MbPos, s->P[0].x, sizeof(s->MbPos));
        memcpy(s->MbVel, s->P[0].v, sizeof(s->MbVel));
        break;
      }
    }
  }
  average_kinematics(c, &ESnap, &s->SpecPot, &s->SpecKin, s->SpecAM, s->Nbound, RefPos, RefVel);
  s->nE = s->Nbound;
  s->Energies = realloc(s->Energies, sizeof(float) * s->Nbound);
  for (HBTInt i = 0; i < s->Nbound; i++) s->Energies[i] = Elist[i].E;
  free(Elist);
  tree_free(&tree);
}

static void sub_append(Sub *s, const Particle *src, int64_t n)
{
  if (s->n + n > s->cap)
  {
    s->cap = (s->n + n) * 2;
    s->P = realloc(s->P, sizeof(Particle) * s->cap);
  }
  memcpy(s->P + s->n, src, sizeof(Particle) * n);
  s->n += n;
}

static void recursive_unbind(const Config *c, Sub *subs, Sub *s)
{ /* Subhalo_t::RecursiveUnbind, src/subhalo_unbind.cpp:432-447 */
  int is_orphan = (s->Nbound <= 1);
  Particle *backup = NULL;
  int64_t nbackup = 0;
  if (is_orphan)
  {
    nbackup = s->n;
    backup = malloc(sizeof(Particle) * (nbackup > 0 ? nbackup : 1));
    memcpy(backup, s->P, sizeof(Particle) * nbackup);
  }
  for (int64_t i = 0; i < s->nnest; i++)
  {
    Sub *child = &subs[s->nest[i]];
    recursive_unbind(c, subs, child);
    sub_append(s, child->P + child->Nbound, child->n - child->Nbound);
  }
  if (is_orphan)
  { /* swap: unbind the single particle, keep the extended list to feed the host */
    Particle *ext = s->P;
    int64_t next = s->n, capext = s->cap;
    s->P = backup;
    s->n = nbackup;
    s->cap = nbackup > 0 ? nbackup : 1;
    unbind(c, s);
    free(s->P);
    s->P = ext;
    s->n = next;
    s->cap = capext;
  }
  else
    unbind(c, s);
}

static void truncate_source(const Config *c, Sub *s)
{ /* Subhalo_t::TruncateSource, src/subhalo_unbind.cpp:449-458 */
  HBTInt Nsource;
  if (s->Nbound <= 1)
    Nsource = s->Nbound;
  else
    Nsource = s->Nbound * c->SourceSubRelaxFactor; /* int*float -> float -> int */
  if (Nsource > s->n) Nsource = (HBTInt)s->n;
  s->n = Nsource;
}

/* ---------------------------------------------------------------------------------------------
 * C entry points (same contract as hbtu_* in include/hbt_unbind.h, ctx -> params)
 * ------------------------------------------------------------------------------------------- */
void hbto_set_num_threads(int n)
{
  omp_set_num_threads(n);
  omp_set_max_active_levels(1);
}
int hbto_get_max_threads(void) { return omp_get_max_threads(); }
int64_t hbto_last_interactions(void) { return g_interactions; }
int64_t hbto_last_opened(void) { return g_opened; }
void hbto_set_shuffle_mode(int mode) { g_shuffle_mode = mode; }
void hbto_seed(unsigned s)
{
  srand(s);
  srand48(s);
}

int hbto_unbind_batch(const hbtu_params *params, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset,
                      const float *pos_mass, const float *vel, const int64_t *nest_offset, const int32_t *nest_list,
                      hbtu_sub_io *io, int32_t flags, int64_t order_capacity, int64_t *order_offset, int32_t *order_out,
                      float *energy_out)
{
  if (params->real_bytes != 4) return HBTU_ERR_UNSUPPORTED;
  Config c;
  config_from(&c, params, epoch);
  c.NoStripping = (flags & HBTU_FLAG_NO_STRIPPING) != 0;
  c.ThermalEnergy = (flags & HBTU_FLAG_THERMAL_ENERGY) != 0;
  g_interactions = g_opened = 0;
  Sub *subs = calloc(nsub > 0 ? nsub : 1, sizeof(Sub));
  char *is_child = calloc(nsub > 0 ? nsub : 1, 1);
  for (int64_t s = 0; s < nsub; s++)
  {
    Sub *sub = &subs[s];
    int64_t b = part_offset[s], n = part_offset[s + 1] - b;
    sub->index = s;
    sub->n = n;
    sub->cap = n > 0 ? n : 1;
    sub->P = malloc(sizeof(Particle) * sub->cap);
    for (int64_t i = 0; i < n; i++)
    {
      Particle *p = &sub->P[i];
      p->Id = (HBTInt)(b + i);
      for (int j = 0; j < 3; j++)
      {
        p->x[j] = pos_mass[4 * (b + i) + j];
        p->v[j] = vel[4 * (b + i) + j];
        p->u = vel[4 * (b + i) + 3];
      }
      p->m = pos_mass[4 * (b + i) + 3];
    }
    for (int j = 0; j < 3; j++)
    {
      sub->AvgPos[j] = io[s].avg_pos[j];
      sub->AvgVel[j] = io[s].avg_vel[j];
      sub->MbPos[j] = io[s].mostbound_pos[j];
      sub->MbVel[j] = io[s].mostbound_vel[j];
    }
    sub->Nbound = (HBTInt)io[s].nbound;
    sub->SinkTrackId = (HBTInt)io[s].sink_track_id;
    sub->Death = io[s].snapshot_index_of_death;
    sub->Sink = io[s].snapshot_index_of_sink;
    sub->SpecPot = io[s].specific_self_potential_energy;
    sub->SpecKin = io[s].specific_self_kinetic_energy;
    for (int j = 0; j < 3; j++) sub->SpecAM[j] = io[s].specific_angular_momentum[j];
    if (nest_offset)
    {
      sub->nest = nest_list + nest_offset[s];
      sub->nnest = nest_offset[s + 1] - nest_offset[s];
    }
  }
  if (nest_offset)
    for (int64_t s = 0; s < nsub; s++)
      for (int64_t k = nest_offset[s]; k < nest_offset[s + 1]; k++)
      {
        int32_t ch = nest_list[k];
        if (ch < 0 || ch >= nsub || is_child[ch] || ch == s) return HBTU_ERR_INVALID;
        is_child[ch] = 1;
      }
  for (int64_t s = 0; s < nsub; s++)
    if ((io[s].flags & HBTU_SUB_PLAIN_UNBIND) && (is_child[s] || subs[s].nnest > 0)) return HBTU_ERR_INVALID;
  for (int64_t s = 0; s < nsub; s++)
    if (!is_child[s])
    { /* field / new-born subhaloes and the merge path call plain Unbind: no orphan rule (src/subhalo_unbind.cpp:498-510,
         src/subhalo_merge.cpp:210) */
      if (io[s].flags & HBTU_SUB_PLAIN_UNBIND) unbind(&c, &subs[s]);
      else recursive_unbind(&c, subs, &subs[s]);
    }
  int rc = HBTU_OK;
  int64_t pos = 0;
  for (int64_t s = 0; s < nsub; s++)
  {
    Sub *sub = &subs[s];
    io[s].nsource_full = sub->n;
    if (flags & HBTU_FLAG_TRUNCATE_SOURCE) truncate_source(&c, sub);
    if (pos + sub->n > order_capacity)
    {
      rc = HBTU_ERR_CAPACITY;
      break;
    }
    order_offset[s] = pos;
    for (int64_t i = 0; i < sub->n; i++) order_out[pos + i] = sub->P[i].Id;
    if (energy_out)
      for (int64_t i = 0; i < sub->n; i++) energy_out[pos + i] = (i < sub->nE) ? sub->Energies[i] : 0.f;
    for (int j = 0; j < 3; j++)
    {
      io[s].avg_pos[j] = sub->AvgPos[j];
      io[s].avg_vel[j] = sub->AvgVel[j];
      io[s].mostbound_pos[j] = sub->MbPos[j];
      io[s].mostbound_vel[j] = sub->MbVel[j];
      io[s].specific_angular_momentum[j] = sub->SpecAM[j];
    }
    io[s].nbound = sub->Nbound;
    io[s].sink_track_id = sub->SinkTrackId;
    io[s].snapshot_index_of_death = sub->Death;
    io[s].snapshot_index_of_sink = sub->Sink;
    io[s].mbound = sub->Mbound;
    io[s].specific_self_potential_energy = sub->SpecPot;
    io[s].specific_self_kinetic_energy = sub->SpecKin;
    io[s].nsource = sub->n;
    io[s].iterations = sub->iterations;
    pos += sub->n;
  }
  if (rc == HBTU_OK) order_offset[nsub] = pos;
  for (int64_t s = 0; s < nsub; s++)
  {
    free(subs[s].P);
    free(subs[s].Energies);
  }
  free(subs);
  free(is_child);
  return rc;
}

static Particle *particles_from(int64_t n, const float *pos_mass)
{
  Particle *P = malloc(sizeof(Particle) * (n > 0 ? n : 1));
  for (int64_t i = 0; i < n; i++)
  {
    P[i].Id = (HBTInt)i;
    for (int j = 0; j < 3; j++)
    {
      P[i].x[j] = pos_mass[4 * i + j];
      P[i].v[j] = 0;
    }
    P[i].m = pos_mass[4 * i + 3];
    P[i].u = 0;
  }
  return P;
}

int hbto_tree_potential(const hbtu_params *params, const hbtu_epoch *epoch, int64_t nsrc, const float *src_pos_mass, int64_t ntgt,
                        const float *tgt_pos, const float *tgt_self_mass, const float *tgt_vel, const double *ref_pos,
                        const double *ref_vel, double *out)
{
  if (params->real_bytes != 4) return HBTU_ERR_UNSUPPORTED;
  if (nsrc < 1) return HBTU_ERR_INVALID;
  Config c;
  config_from(&c, params, epoch);
  Particle *P = particles_from(nsrc, src_pos_mass);
  View view = {P, NULL, 1.f};
  Tree tree;
  memset(&tree, 0, sizeof(tree));
  tree_build(&tree, &c, view, (HBTInt)nsrc);
  HBTReal rp[3] = {0, 0, 0}, rv[3] = {0, 0, 0};
  if (tgt_vel)
    for (int j = 0; j < 3; j++)
    {
      rp[j] = ref_pos[j];
      rv[j] = ref_vel[j];
    }
  int64_t cnt0 = 0, cnt1 = 0;
#pragma omp parallel for reduction(+ : cnt0, cnt1) schedule(dynamic, 256)
  for (int64_t i = 0; i < ntgt; i++)
  {
    int64_t cnt[2] = {0, 0};
    HBTReal m = tgt_self_mass ? tgt_self_mass[i] : 0.f;
    if (tgt_vel)
      out[i] = binding_energy(&tree, &tgt_pos[4 * i], &tgt_vel[4 * i], rp, rv, m, cnt);
    else
      out[i] = tree_potential(&tree, &tgt_pos[4 * i], m, cnt);
    cnt0 += cnt[0];
    cnt1 += cnt[1];
  }
  g_interactions = cnt0;
  g_opened = cnt1;
  tree_free(&tree);
  free(P);
  return HBTU_OK;
}

/* Instrumented walk: per-target number of accepted sources and of opened nodes (the algorithmic
 * work unit of SURVEY.md section 8(d)); also returns the number of tree cells via the return value. */
int hbto_walk_counts(const hbtu_params *params, const hbtu_epoch *epoch, int64_t nsrc, const float *src_pos_mass, int64_t ntgt,
                     const float *tgt_pos, int64_t *accepted, int64_t *opened)
{
  if (nsrc < 1) return HBTU_ERR_INVALID;
  Config c;
  config_from(&c, params, epoch);
  Particle *P = particles_from(nsrc, src_pos_mass);
  View view = {P, NULL, 1.f};
  Tree tree;
  memset(&tree, 0, sizeof(tree));
  int ncell = tree_build(&tree, &c, view, (HBTInt)nsrc);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < ntgt; i++)
  {
    int64_t cnt[2] = {0, 0};
    tree_potential(&tree, &tgt_pos[4 * i], 0.f, cnt);
    if (accepted) accepted[i] = cnt[0];
    if (opened) opened[i] = cnt[1];
  }
  tree_free(&tree);
  free(P);
  return ncell;
}

/* ---------------------------------------------------------------------------------------------
 * Post-unbinding properties: Subhalo_t::CalculateProfileProperties
 * (src/subhalo.cpp:242-332) and Subhalo_t::CalculateShape (src/subhalo.cpp:334-398).
 * ------------------------------------------------------------------------------------------- */
typedef struct
{ /* RadMassVel_t, src/snapshot.h:41-49 */
  HBTReal r, m, v;
} RadMassVel;

static int comp_prof_radius(const void *a, const void *b)
{ /* CompProfRadius, src/subhalo.cpp:233-236 (std::sort is unstable: ties carry equal r) */
  const HBTReal x = ((const RadMassVel *)a)->r, y = ((const RadMassVel *)b)->r;
  return (x > y) - (x < y);
}

static HBTReal periodic_distance(const Config *c, const HBTReal x[3], const HBTReal y[3])
{ /* PeriodicDistance, src/config_parser.h:143-156: HBTReal differences, NEAREST, sqrt of the HBTReal sum */
  HBTReal dx[3];
  for (int j = 0; j < 3; j++)
  {
    dx[j] = x[j] - y[j];
    if (c->Periodic) dx[j] = nearest_f(c, dx[j]);
  }
  return sqrtf(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]);
}

static void profile_properties(const Config *c, const float *pm, hbtu_profile_io *o)
{ /* src/subhalo.cpp:242-332; pm = the subhalo's particle list (x,y,z,m per particle) */
  const HBTInt Nbound = (HBTInt)o->nbound;
  if (Nbound <= 1)
  { /* :265-286 */
    o->rmax_comoving = 0.f;
    o->vmax_physical = 0.f;
    o->r2sigma_comoving = 0.f;
    o->rhalf_comoving = 0.f;
    o->bound_r200crit_comoving = 0.f;
    o->bound_m200crit = 0.f;
    return;
  }
  const HBTReal VelocityUnit = c->G / c->ScaleFactor; /* :287 */
  const HBTReal cen[3] = {(HBTReal)o->mostbound_pos[0], (HBTReal)o->mostbound_pos[1], (HBTReal)o->mostbound_pos[2]};
  RadMassVel *prof = (RadMassVel *)malloc(sizeof(RadMassVel) * (size_t)Nbound);
  for (HBTInt i = 0; i < Nbound; i++)
  { /* :290-294 */
    prof[i].r = periodic_distance(c, cen, &pm[4 * (size_t)i]);
    prof[i].m = pm[4 * (size_t)i + 3];
  }
  qsort(prof, (size_t)Nbound, sizeof(RadMassVel), comp_prof_radius); /* :297 */
  double m_cum = 0.;
  for (HBTInt i = 0; i < Nbound; i++) prof[i].m = (HBTReal)(m_cum += prof[i].m); /* :298-299 */
  for (HBTInt i = 0; i < Nbound; i++)
  { /* :302-306 */
    if (prof[i].r < c->SofteningHalo) prof[i].r = c->SofteningHalo;
    prof[i].v = prof[i].m / prof[i].r;
  }
  HBTInt imax = 0; /* max_element: first of the largest, :308 */
  for (HBTInt i = 1; i < Nbound; i++)
    if (prof[imax].v < prof[i].v) imax = i;
  o->rmax_comoving = prof[imax].r;
  o->vmax_physical = sqrtf(prof[imax].v * VelocityUnit);
  o->rhalf_comoving = prof[Nbound / 2].r;
  o->r2sigma_comoving = prof[(HBTInt)(Nbound * 0.955)].r;
  /* HaloVirialFactors: virialF_c200 = 200 (src/snapshot.cpp:334); SphericalOverdensitySize(prof), :264-281 */
  const HBTReal VirialFactor = 200.f;
  const HBTReal RhoVirial =
      (HBTReal)(VirialFactor * c->Hz * c->Hz / 2.0 / c->G * c->ScaleFactor * c->ScaleFactor * c->ScaleFactor);
  for (HBTInt i = Nbound - 1; i >= 0; i--)
  {
    const HBTReal r = prof[i].r, m = prof[i].m;
    if (m > RhoVirial * r * r * r)
    {
      o->bound_m200crit = m;
      o->bound_r200crit_comoving = (float)pow(m / RhoVirial, 1.0 / 3);
      break;
    }
  }
  if (o->vmax_physical >= o->last_max_vmax_physical)
  { /* :322-326 */
    o->snapshot_index_of_last_max_vmax = c->SnapshotIndex;
    o->last_max_vmax_physical = o->vmax_physical;
  }
  free(prof);
}

static void shape(const Config *c, const float *pm, hbtu_profile_io *o)
{ /* src/subhalo.cpp:334-398 (without the HAS_GSL eigen-vectors) */
  const HBTInt Nbound = (HBTInt)o->nbound;
  if (Nbound <= 1)
  {
    for (int j = 0; j < 6; j++) o->inertial_tensor[j] = o->inertial_tensor_weighted[j] = 0.f;
    return;
  }
  const HBTReal cen[3] = {(HBTReal)o->mostbound_pos[0], (HBTReal)o->mostbound_pos[1], (HBTReal)o->mostbound_pos[2]};
  double I[6] = {0, 0, 0, 0, 0, 0}, Iw[6] = {0, 0, 0, 0, 0, 0}; /* xx, xy, xz, yy, yz, zz */
  for (HBTInt i = 1; i < Nbound; i++)
  {
    const HBTReal m = pm[4 * (size_t)i + 3];
    HBTReal dx = pm[4 * (size_t)i] - cen[0], dy = pm[4 * (size_t)i + 1] - cen[1], dz = pm[4 * (size_t)i + 2] - cen[2];
    if (c->Periodic)
    {
      dx = nearest_f(c, dx);
      dy = nearest_f(c, dy);
      dz = nearest_f(c, dz);
    }
    const HBTReal dx2 = dx * dx, dy2 = dy * dy, dz2 = dz * dz;
    I[0] += dx2 * m; I[3] += dy2 * m; I[5] += dz2 * m;
    I[1] += dx * dy * m; I[2] += dx * dz * m; I[4] += dy * dz * m;
    HBTReal dr2 = dx2 + dy2 + dz2;
    dr2 /= m;
    Iw[0] += dx2 / dr2; Iw[3] += dy2 / dr2; Iw[5] += dz2 / dr2;
    Iw[1] += dx * dy / dr2; Iw[2] += dx * dz / dr2; Iw[4] += dy * dz / dr2;
  }
  for (int j = 0; j < 6; j++)
  { /* assigned to float, then /= Mbound (float), :389-392 */
    float a = (float)I[j], b = (float)Iw[j];
    a /= o->mbound;
    b /= o->mbound;
    o->inertial_tensor[j] = a;
    o->inertial_tensor_weighted[j] = b;
  }
}

int hbto_profile_batch(const hbtu_params *params, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset,
                       const float *pos_mass, hbtu_profile_io *io)
{
  Config c;
  config_from(&c, params, epoch);
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t s = 0; s < nsub; s++)
  {
    if (io[s].nbound > part_offset[s + 1] - part_offset[s]) continue; /* caller error; the GPU entry point rejects it */
    profile_properties(&c, &pos_mass[4 * part_offset[s]], &io[s]);
    shape(&c, &pos_mass[4 * part_offset[s]], &io[s]);
  }
  return HBTU_OK;
}


/* ---------------------------------------------------------------------------------------------
 * Source preparation: SubhaloSnapshot_t::MaskSubhalos + SubhaloMasker_t::Mask,
 * src/subhalo_tracking.cpp:793-841.
 * ------------------------------------------------------------------------------------------- */
typedef struct
{ /* unordered_set<HBTInt> ExclusionList (:795): open addressing, insert-only */
  int64_t *key;
  unsigned char *used;
  uint64_t mask;
} IdSet;

static int idset_insert(IdSet *h, int64_t id)
{ /* returns 1 when inserted (was absent), like unordered_set::insert(...).second (:812-813) */
  uint64_t z = (uint64_t)id * 0x9E3779B97F4A7C15ULL;
  uint64_t i = (z ^ (z >> 29)) & h->mask;
  while (h->used[i])
  {
    if (h->key[i] == id) return 0;
    i = (i + 1) & h->mask;
  }
  h->used[i] = 1;
  h->key[i] = id;
  return 1;
}

static void mask_recursive(int64_t s, const int64_t *part_offset, const int64_t *ids, const int64_t *nest_offset,
                           const int32_t *nest_list, const int64_t *nbound, IdSet *h, int64_t *new_count, int32_t *keep_index)
{ /* SubhaloMasker_t::Mask, :801-822 */
  if (nest_offset)
    for (int64_t k = nest_offset[s]; k < nest_offset[s + 1]; k++)
      mask_recursive(nest_list[k], part_offset, ids, nest_offset, nest_list, nbound, h, new_count, keep_index);
  const int64_t b = part_offset[s], e = part_offset[s + 1];
  int64_t save = b;
  if (nbound[s] <= 1)
  { /* skip orphans (:806): list untouched, nothing excluded */
    for (int64_t i = b; i < e; i++) keep_index[save++] = (int32_t)i;
  }
  else
    for (int64_t i = b; i < e; i++)
      if (idset_insert(h, ids[i])) keep_index[save++] = (int32_t)i;
  new_count[s] = save - b;
}

static int64_t tree_particles(int64_t s, const int64_t *part_offset, const int64_t *nest_offset, const int32_t *nest_list)
{
  int64_t n = part_offset[s + 1] - part_offset[s];
  if (nest_offset)
    for (int64_t k = nest_offset[s]; k < nest_offset[s + 1]; k++) n += tree_particles(nest_list[k], part_offset, nest_offset, nest_list);
  return n;
}

int hbto_mask_batch(const hbtu_params *params, int64_t nsub, const int64_t *part_offset, const int64_t *particle_id,
                    const int64_t *nest_offset, const int32_t *nest_list, const int64_t *nbound, int64_t *new_count,
                    int32_t *keep_index)
{
  (void)params;
  char *is_child = calloc((size_t)(nsub > 0 ? nsub : 1), 1);
  if (nest_offset)
    for (int64_t k = 0; k < nest_offset[nsub]; k++)
    {
      if (nest_list[k] < 0 || nest_list[k] >= nsub || is_child[nest_list[k]]) { free(is_child); return HBTU_ERR_INVALID; }
      is_child[nest_list[k]] = 1;
    }
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t s = 0; s < nsub; s++)
  { /* one SubhaloMasker_t per host group = per root, :834 */
    if (is_child[s]) continue;
    int64_t n = tree_particles(s, part_offset, nest_offset, nest_list);
    uint64_t cap = 16;
    while (cap < 2 * (uint64_t)n + 2) cap <<= 1;
    IdSet h = {malloc(sizeof(int64_t) * cap), calloc(cap, 1), cap - 1};
    mask_recursive(s, part_offset, particle_id, nest_offset, nest_list, nbound, &h, new_count, keep_index);
    free(h.key);
    free(h.used);
  }
  free(is_child);
  return HBTU_OK;
}


/* ---------------------------------------------------------------------------------------------
 * Particle query: MappedIndexTable_t::Fill (src/hash.tpp:18-32) +
 * MappedIndexTable_t::GetIndices (src/hash_remote.tpp:9-88).
 * ------------------------------------------------------------------------------------------- */
typedef struct
{ /* IndexedKey_t, src/hash.h */
  int64_t Key, Index;
} IdPair;

static int comp_pair(const void *a, const void *b)
{ /* CompPair, src/hash.tpp:13-17; the index breaks ties so that the answer for duplicated Ids is defined */
  const IdPair *x = (const IdPair *)a, *y = (const IdPair *)b;
  if (x->Key != y->Key) return (x->Key > y->Key) - (x->Key < y->Key);
  return (x->Index > y->Index) - (x->Index < y->Index);
}

int hbto_idtable_query(const hbtu_params *params, int64_t n, const int64_t *particle_id, int64_t nq, const int64_t *query_id,
                       int64_t *index_out)
{
  (void)params;
  IdPair *Map = malloc(sizeof(IdPair) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; i++)
  {
    Map[i].Key = particle_id[i];
    Map[i].Index = i;
  }
  qsort(Map, (size_t)n, sizeof(IdPair), comp_pair); /* Fill, :29 */
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < nq; q++)
  { /* lower_bound + equality test, src/hash_remote.tpp:76-83 */
    const int64_t key = query_id[q];
    int64_t lo = 0, hi = n;
    while (lo < hi)
    {
      int64_t mid = lo + ((hi - lo) >> 1);
      if (Map[mid].Key < key) lo = mid + 1; else hi = mid;
    }
    index_out[q] = (lo < n && Map[lo].Key == key) ? Map[lo].Index : -1;
  }
  free(Map);
  return HBTU_OK;
}


/* ---------------------------------------------------------------------------------------------
 * Merger trap detection: src/subhalo_merge.cpp:29-172.
 * ------------------------------------------------------------------------------------------- */
#define NUM_PART_CORE_MAX 20
#define DELTA_CRIT 2.

typedef struct
{ /* SubHelper_t, src/subhalo_merge.cpp:14-27 */
  int64_t HostTrackId;
  int IsMerged;
  HBTReal ComovingPosition[3], PhysicalVelocity[3];
  float ComovingSigmaR, PhysicalSigmaV;
} SubHelper;

static void helper_build(const Config *c, SubHelper *h, int64_t nbound, const float *pm, const float *vel)
{ /* BuildPosition :29-80, BuildVelocity :81-123 */
  if (nbound == 0) { h->ComovingSigmaR = 0.f; h->PhysicalSigmaV = 0.f; return; }
  if (nbound == 1)
  {
    h->ComovingSigmaR = 0.f; h->PhysicalSigmaV = 0.f;
    for (int j = 0; j < 3; j++) { h->ComovingPosition[j] = pm[j]; h->PhysicalVelocity[j] = vel[j]; }
    return;
  }
  const int64_t NumPart = nbound > NUM_PART_CORE_MAX ? NUM_PART_CORE_MAX : nbound;
  double sx[3] = {0, 0, 0}, sx2[3] = {0, 0, 0}, sv[3] = {0, 0, 0}, sv2[3] = {0, 0, 0}, origin[3] = {0, 0, 0}, msum = 0.;
  if (c->Periodic)
    for (int j = 0; j < 3; j++) origin[j] = pm[j];
  for (int64_t i = 0; i < NumPart; i++)
  {
    const HBTReal m = pm[4 * i + 3];
    msum += m;
    for (int j = 0; j < 3; j++)
    {
      double dx = c->Periodic ? nearest_d(c, pm[4 * i + j] - origin[j]) : pm[4 * i + j];
      sx[j] += dx * m;
      sx2[j] += dx * dx * m;
      double dv = vel[4 * i + j];
      sv[j] += dv * m;
      sv2[j] += dv * dv * m;
    }
  }
  for (int j = 0; j < 3; j++)
  {
    sx[j] /= msum; sx2[j] /= msum;
    h->ComovingPosition[j] = (HBTReal)sx[j];
    if (c->Periodic) h->ComovingPosition[j] += origin[j];
    sx2[j] -= sx[j] * sx[j];
    sv[j] /= msum; sv2[j] /= msum;
    h->PhysicalVelocity[j] = (HBTReal)sv[j];
    sv2[j] -= sv[j] * sv[j];
  }
  h->ComovingSigmaR = (float)sqrt(sx2[0] + sx2[1] + sx2[2]);
  h->PhysicalSigmaV = (float)sqrt(sv2[0] + sv2[1] + sv2[2]);
}

static float sink_distance(const Config *c, const hbtu_trap_io *sat, const SubHelper *cen)
{ /* SinkDistance, :125-130 */
  const HBTReal sp[3] = {(HBTReal)sat->mostbound_pos[0], (HBTReal)sat->mostbound_pos[1], (HBTReal)sat->mostbound_pos[2]};
  const HBTReal sv[3] = {(HBTReal)sat->mostbound_vel[0], (HBTReal)sat->mostbound_vel[1], (HBTReal)sat->mostbound_vel[2]};
  float d = periodic_distance(c, cen->ComovingPosition, sp);
  HBTReal dv[3] = {cen->PhysicalVelocity[0] - sv[0], cen->PhysicalVelocity[1] - sv[1], cen->PhysicalVelocity[2] - sv[2]};
  float v = sqrtf(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]);
  return d / cen->ComovingSigmaR + v / cen->PhysicalSigmaV;
}

int hbto_detect_traps(const hbtu_params *params, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset, const float *pos_mass,
                      const float *vel, const int64_t *nest_offset, const int32_t *nest_list, hbtu_trap_io *io)
{
  Config c;
  config_from(&c, params, epoch);
  SubHelper *H = calloc((size_t)(nsub > 0 ? nsub : 1), sizeof(SubHelper));
  for (int64_t i = 0; i < nsub; i++) H[i].HostTrackId = -1;
  if (nest_offset) /* FillHostTrackIds, :163-172 */
    for (int64_t i = 0; i < nsub; i++)
      for (int64_t k = nest_offset[i]; k < nest_offset[i + 1]; k++) H[nest_list[k]].HostTrackId = i;
  for (int64_t i = 0; i < nsub; i++) helper_build(&c, &H[i], io[i].nbound, &pos_mass[4 * part_offset[i]], &vel[4 * part_offset[i]]);
  for (int64_t i = 0; i < nsub; i++)
  { /* DetectTraps, :132-161 */
    if (io[i].sink_track_id != -1) continue;
    int64_t HostId = H[i].HostTrackId;
    while (HostId >= 0)
    {
      if (io[HostId].nbound > 1)
      {
        float delta = sink_distance(&c, &io[i], &H[HostId]);
        if (delta < DELTA_CRIT)
        {
          io[i].sink_track_id = HostId;
          io[i].snapshot_index_of_sink = c.SnapshotIndex;
          if (io[i].nbound > 1) H[HostId].IsMerged = 1;
          break;
        }
      }
      HostId = H[HostId].HostTrackId;
    }
  }
  for (int64_t i = 0; i < nsub; i++) io[i].is_merged = H[i].IsMerged;
  free(H);
  return HBTU_OK;
}
