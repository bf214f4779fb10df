/* polb200 oracle instrumentation (our code, compiled only into the scratch copy of the
   reference).  When the environment variable POLB200_DUMP is set to a path prefix, every call
   of PairLJCutCoulLongPolarization::compute() writes <prefix>.<call>.bin holding the inputs and
   outputs of that call as named records:  name[16] | dtype 'i'(int32) or 'd'(float64) | int64 n | data */
#ifndef POLB200_DUMP_H
#define POLB200_DUMP_H
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "atom.h"
#include "domain.h"
#include "force.h"
#include "neigh_list.h"

namespace polb200_dump {
static std::vector<double> mu_in;
static int ncall = 0;

static inline void rec(FILE *fp, const char *name, char dtype, long long n, const void *data)
{
  char nm[16];
  memset(nm,0,16);
  strncpy(nm,name,15);
  fwrite(nm,1,16,fp);
  fwrite(&dtype,1,1,fp);
  fwrite(&n,sizeof(long long),1,fp);
  fwrite(data,(dtype == 'i') ? 4 : 8,n,fp);
}
static inline void reci(FILE *fp, const char *name, int v) { rec(fp,name,'i',1,&v); }
static inline void recd(FILE *fp, const char *name, double v) { rec(fp,name,'d',1,&v); }

static inline void pre(LAMMPS_NS::Atom *atom)
{
  if (!getenv("POLB200_DUMP")) return;
  int n = atom->nlocal;
  mu_in.resize(3*(size_t)n);
  for (int i = 0; i < n; i++)
    for (int p = 0; p < 3; p++) mu_in[3*i+p] = atom->mu_induced[i][p];
}

static inline void post(LAMMPS_NS::Atom *atom, LAMMPS_NS::Domain *domain, LAMMPS_NS::Force *force,
                        LAMMPS_NS::NeighList *list, int eflag, int vflag, int iterations,
                        double eng_vdwl, double eng_coul, double eng_pol, const double *virial,
                        double g_ewald, double cut_coul, double tabinnersq,
                        int ncoultablebits, int ncoulmask, int ncoulshiftbits,
                        const double *rtable, const double *drtable, const double *ftable,
                        const double *dftable, const double *ctable, const double *dctable,
                        const double *etable, const double *detable,
                        double **cutsq, double **cut_ljsq, double **lj1, double **lj2,
                        double **lj3, double **lj4, double **offset, int neighbor_ago)
{
  const char *prefix = getenv("POLB200_DUMP");
  if (!prefix) return;
  const char *maxs = getenv("POLB200_DUMP_MAX");
  int maxcall = maxs ? atoi(maxs) : 1000000;
  if (ncall >= maxcall) { ncall++; return; }
  char fname[1024];
  snprintf(fname,1024,"%s.%d.bin",prefix,ncall++);
  FILE *fp = fopen(fname,"wb");
  if (!fp) return;
  int nlocal = atom->nlocal, nghost = atom->nghost, nall = nlocal+nghost, nt = atom->ntypes;
  reci(fp,"nlocal",nlocal); reci(fp,"nghost",nghost); reci(fp,"ntypes",nt);
  reci(fp,"eflag",eflag); reci(fp,"vflag",vflag); reci(fp,"iterations",iterations);
  reci(fp,"newton_pair",force->newton_pair);
  reci(fp,"neighbor_ago",neighbor_ago);   /* Neighbor::ago: 0 on the steps the lists were rebuilt (src/neighbor.cpp:1923-1937) */
  reci(fp,"ncoultablebits",ncoultablebits); reci(fp,"ncoulmask",ncoulmask);
  reci(fp,"ncoulshiftbits",ncoulshiftbits);
  rec(fp,"boxlo",'d',3,domain->boxlo); rec(fp,"boxhi",'d',3,domain->boxhi);
  recd(fp,"eng_vdwl",eng_vdwl); recd(fp,"eng_coul",eng_coul); recd(fp,"eng_pol",eng_pol);
  rec(fp,"virial",'d',6,virial);
  recd(fp,"g_ewald",g_ewald); recd(fp,"cut_coul",cut_coul); recd(fp,"tabinnersq",tabinnersq);
  recd(fp,"qqrd2e",force->qqrd2e);
  rec(fp,"special_lj",'d',4,force->special_lj); rec(fp,"special_coul",'d',4,force->special_coul);
  rec(fp,"tag",'i',nall,atom->tag); rec(fp,"type",'i',nall,atom->type);
  rec(fp,"molecule",'i',nall,atom->molecule);
  rec(fp,"x",'d',3LL*nall,&atom->x[0][0]); rec(fp,"q",'d',nall,atom->q);
  rec(fp,"alpha",'d',nall,atom->static_polarizability);
  rec(fp,"mu_in",'d',3LL*nlocal,mu_in.data());
  rec(fp,"mu_out",'d',3LL*nlocal,&atom->mu_induced[0][0]);
  rec(fp,"ef_static",'d',3LL*nlocal,&atom->ef_static[0][0]);
  if (eflag / 2) rec(fp,"eatom",'d',(long long)nall,force->pair->eatom);
  if (vflag / 4) rec(fp,"vatom",'d',6LL*nall,&force->pair->vatom[0][0]);
  rec(fp,"f",'d',3LL*nall,&atom->f[0][0]);
  int nn = (nt+1)*(nt+1);
  rec(fp,"cutsq",'d',nn,&cutsq[0][0]); rec(fp,"cut_ljsq",'d',nn,&cut_ljsq[0][0]);
  rec(fp,"lj1",'d',nn,&lj1[0][0]); rec(fp,"lj2",'d',nn,&lj2[0][0]);
  rec(fp,"lj3",'d',nn,&lj3[0][0]); rec(fp,"lj4",'d',nn,&lj4[0][0]);
  rec(fp,"offset",'d',nn,&offset[0][0]);
  if (ncoultablebits && ncall == 1) {
    int ntable = 1 << ncoultablebits;
    rec(fp,"rtable",'d',ntable,rtable); rec(fp,"drtable",'d',ntable,drtable);
    rec(fp,"ftable",'d',ntable,ftable); rec(fp,"dftable",'d',ntable,dftable);
    rec(fp,"ctable",'d',ntable,ctable); rec(fp,"dctable",'d',ntable,dctable);
    rec(fp,"etable",'d',ntable,etable); rec(fp,"detable",'d',ntable,detable);
  }
  int inum = list->inum;
  reci(fp,"inum",inum);
  rec(fp,"ilist",'i',inum,list->ilist);
  std::vector<int> nn_i(inum), flat;
  for (int ii = 0; ii < inum; ii++) {
    int i = list->ilist[ii];
    nn_i[ii] = list->numneigh[i];
    flat.insert(flat.end(),list->firstneigh[i],list->firstneigh[i]+list->numneigh[i]);
  }
  rec(fp,"numneigh",'i',inum,nn_i.data());
  rec(fp,"neigh",'i',(long long)flat.size(),flat.data());
  if (atom->nspecial && atom->maxspecial > 0) {
    reci(fp,"maxspecial",atom->maxspecial);
    rec(fp,"nspecial",'i',3LL*nlocal,&atom->nspecial[0][0]);
    rec(fp,"special",'i',(long long)nlocal*atom->maxspecial,&atom->special[0][0]);
  }
  fclose(fp);
}
}
#endif
