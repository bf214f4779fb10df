#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the repaired reference binary (oracle/_ref/lmp_serial).

Runs HERE (needs /root/reference for the example inputs); the GPU box only sees the fixtures.
Each fixture = inputs and outputs of ONE PairLJCutCoulLongPolarization::compute() call of the
reference, dumped by the hooks of oracle/ref_shims/polb200_dump.h, keyed by local atom index
(ghost forces folded onto their owners the way Verlet's reverse_comm does).

Also transcribes the thermo tables of the reference's committed logs into thermo_logs.json.
"""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import polref as P  # noqa: E402

REF = Path(os.environ.get("POLB200_REFERENCE", "/root/reference"))
EX = REF / "polarization" / "examples"
LMP = ROOT / "oracle" / "_ref" / "lmp_serial"
OUT = ROOT / "tests" / "golden"

H2_STYLE = ("pair_style lj/cut/coul/long/polarization 2.5 10.797442 precision 0.00000000001 "
            "max_iterations 100 damp_type exponential damp 2.1304 polar_gs_ranked yes debug no use_previous yes")

# (case name, example dir, input file, replacement pair_style line or None, extra lines before run,
#  steps to keep, run length)
CASES = [
    ("h2_default", "Bulk H2", "h2.input", None, [], [0, 1, 2, 3], 3),
    ("h2_jacobi_fixed3", "Bulk H2", "h2.input",
     H2_STYLE.replace("polar_gs_ranked yes", "polar_gs_ranked no fixed_iteration yes").replace(
         "max_iterations 100", "max_iterations 3"), [], [0, 1], 1),
    ("h2_jacobi_fixed30", "Bulk H2", "h2.input",
     H2_STYLE.replace("polar_gs_ranked yes", "polar_gs_ranked no fixed_iteration yes").replace(
         "max_iterations 100", "max_iterations 30"), [], [0], 0),
    ("h2_jacobi_precision", "Bulk H2", "h2.input",
     H2_STYLE.replace("polar_gs_ranked yes", "polar_gs_ranked no"), [], [0, 1], 1),
    ("h2_gs", "Bulk H2", "h2.input",
     H2_STYLE.replace("polar_gs_ranked yes", "polar_gs_ranked no polar_gs yes"), [], [0, 1], 1),
    ("h2_gs_fixed3", "Bulk H2", "h2.input",
     H2_STYLE.replace("polar_gs_ranked yes", "polar_gs_ranked no polar_gs yes fixed_iteration yes").replace(
         "max_iterations 100", "max_iterations 3"), [], [0], 0),
    ("h2_zodid", "Bulk H2", "h2.input",
     H2_STYLE.replace("polar_gs_ranked yes", "polar_gs_ranked no zodid yes").replace("use_previous yes",
                                                                                      "use_previous no"),
     [], [0], 0),
    ("h2_nodamp", "Bulk H2", "h2.input", H2_STYLE.replace("damp_type exponential", "damp_type none"),
     [], [0], 0),
    ("h2_noprev", "Bulk H2", "h2.input", H2_STYLE.replace("use_previous yes", "use_previous no"),
     [], [0, 1], 1),
    ("h2_diverge", "Bulk H2", "h2.input", H2_STYLE.replace("max_iterations 100", "max_iterations 3"),
     [], [0], 0),
    ("h2_notable", "Bulk H2", "h2.input", None, ["pair_modify table 0"], [0], 0),
    # newton_pair off: half/bin/newtoff list, no ghost forces, pairwise virial (vflag = 1)
    ("h2_newtonoff", "Bulk H2", "h2.input",
     H2_STYLE.replace("polar_gs_ranked yes", "polar_gs_ranked no fixed_iteration yes").replace(
         "max_iterations 100", "max_iterations 5"), ["__NEWTON_OFF__"], [0], 0),
    # per-atom energy / virial tallies (Pair::ev_tally, ev_tally_xyz): requested through pe/atom + stress/atom
    ("h2_peratom", "Bulk H2", "h2.input",
     H2_STYLE.replace("polar_gs_ranked yes", "polar_gs_ranked no fixed_iteration yes").replace(
         "max_iterations 100", "max_iterations 5"),
     ["compute pea all pe/atom pair", "compute sta all stress/atom NULL pair", "compute spe all reduce sum c_pea",
      "compute sst all reduce sum c_sta[1] c_sta[2] c_sta[3] c_sta[4] c_sta[5] c_sta[6]",
      "thermo_style custom step pe c_spe c_sst[1] c_sst[2] c_sst[3] c_sst[4] c_sst[5] c_sst[6]"], [0], 0),
    # neigh_modify exclude: the pair loop only sees what the list holds; the polarization loops do not care
    ("h2_exclude_intra", "Bulk H2", "h2.input",
     H2_STYLE.replace("polar_gs_ranked yes", "polar_gs_ranked no fixed_iteration yes").replace(
         "max_iterations 100", "max_iterations 5"), ["neigh_modify exclude molecule/intra all"], [0], 0),
    ("h2_exclude_mixed", "Bulk H2", "h2.input",
     H2_STYLE.replace("polar_gs_ranked yes", "polar_gs_ranked no fixed_iteration yes").replace(
         "max_iterations 100", "max_iterations 5"),
     ["group gsite type 1", "group esite type 2", "neigh_modify exclude type 1 3",
      "neigh_modify exclude group gsite esite exclude molecule/intra esite"], [0], 0),
    ("methane_default", "MOF5+Methane", "MOF5+PCRC.restart.pdb.input", None, [], [0, 1, 2], 2),
    # shipped input aborts in fix rigid; single-point compute() with the integrator swapped
    ("co2_singlepoint", "MOF5+CO2", "co2_mof5.restart.pdb.input", None, ["__NVE__"], [0], 0),
]


def run_case(name, exdir, inp, style, extra, keep, nrun):
    work = Path(tempfile.mkdtemp(prefix=f"polgold_{name}_"))
    src = EX / exdir
    for f in src.iterdir():
        if f.suffix in (".data",) or f.name.endswith(".data"):
            shutil.copy(f, work / f.name)
    text = (src / inp).read_text()
    text = text.replace("ewald/disp", "ewald")
    text = re.sub(r"^dump\S* .*$", "", text, flags=re.M)
    if style is not None:
        text = re.sub(r"^pair_style .*$", style, text, flags=re.M)
    text = re.sub(r"(variable\s+nstep\s+equal\s+)\d+", rf"\g<1>{nrun}", text)
    nve = "__NVE__" in extra
    if "__NEWTON_OFF__" in extra:
        text = re.sub(r"^(atom_style.*)$", r"\1\nnewton off", text, count=1, flags=re.M)
    extra = [e for e in extra if e not in ("__NVE__", "__NEWTON_OFF__")]
    if nve:
        text = re.sub(r"^fix\s+rigid_nve.*$", "fix 1 moving nve", text, flags=re.M)
    lines = text.splitlines()
    idx = max(i for i, l in enumerate(lines) if l.strip().startswith("run"))
    lines[idx:idx] = extra
    (work / "in.case").write_text("\n".join(lines) + "\n")
    env = dict(os.environ, POLB200_DUMP=str(work / "dump"), POLB200_DUMP_MAX=str(max(keep) + 1))
    r = subprocess.run([str(LMP), "-in", "in.case", "-echo", "none"], cwd=work, env=env, capture_output=True,
                       text=True)
    if r.returncode != 0:
        print(r.stdout[-3000:])
        raise SystemExit(f"{name}: lmp_serial failed")
    log = (work / "log.lammps").read_text()
    pair_style = [l for l in lines if l.strip().startswith("pair_style")][0]
    pair_coeffs = [l for l in lines if l.strip().startswith("pair_coeff")]
    pair_modify = [l for l in lines if l.strip().startswith("pair_modify")]
    m = re.search(r"G vector \(1/distance\) = (\S+)", log)
    thermo = parse_thermo(log)
    first = None
    for step in range(max(keep) + 1):
        d = P.read_refdump(work / f"dump.{step}.bin")
        if step == 0:
            first = d
        if step not in keep:
            continue
        nl = int(d["nlocal"][0])
        nt = int(d["ntypes"][0])
        x = d["x"].reshape(-1, 3)
        tag = d["tag"]
        f_all = d["f"].reshape(-1, 3)
        # fold ghost forces onto owners (tag -> local index)
        loc = {int(t): i for i, t in enumerate(tag[:nl])}
        f_own = f_all[:nl].copy()
        for g in range(nl, len(tag)):
            f_own[loc[int(tag[g])]] += f_all[g]
        ms = int(d["maxspecial"][0]) if "maxspecial" in d else 0
        fx = dict(
            x=x[:nl], q=d["q"][:nl], type=d["type"][:nl], molecule=d["molecule"][:nl], alpha=d["alpha"][:nl],
            tag=tag[:nl], boxlo=d["boxlo"], boxhi=d["boxhi"], ntypes=nt,
            mu_in=d["mu_in"].reshape(nl, 3), mu_out=d["mu_out"].reshape(nl, 3),
            ef_static=d["ef_static"].reshape(nl, 3), f=f_own,
            eng_vdwl=d["eng_vdwl"][0], eng_coul=d["eng_coul"][0], eng_pol=d["eng_pol"][0],
            virial=d["virial"], iterations=int(d["iterations"][0]), eflag=int(d["eflag"][0]),
            vflag=int(d["vflag"][0]), g_ewald=d["g_ewald"][0], nghost=int(d["nghost"][0]),
            special_lj=d["special_lj"], special_coul=d["special_coul"],
            numneigh_half=d["numneigh"], npairs_half=len(d["neigh"]),
            pair_style=pair_style, pair_coeff="\n".join(pair_coeffs), pair_modify="\n".join(pair_modify),
            step=step, ncoultablebits=int(d["ncoultablebits"][0]),
        )
        nm = [l for l in lines if l.strip().startswith("neigh_modify") and "exclude" in l]
        if nm:  # exclusion rules + the group masks they refer to (groups of these cases are defined by type)
            bits, mask = {"all": 1}, np.ones(nl, dtype=np.int32)
            for l in lines:
                w = l.split()
                if len(w) == 4 and w[0] == "group" and w[2] == "type":
                    bits[w[1]] = 1 << len(bits)
                    mask[d["type"][:nl] == int(w[3])] |= bits[w[1]]
            fx["neigh_modify"] = "\n".join(nm)
            fx["mask"] = mask
            fx["group_names"] = np.array(list(bits.keys()))
            fx["group_bits"] = np.array(list(bits.values()), dtype=np.int32)
        for key, width in (("eatom", 1), ("vatom", 6)):
            if key in d:  # fold ghost tallies onto their owners (what reverse_comm of the computes does)
                a = d[key].reshape(-1, width)
                own = a[:nl].copy()
                for g in range(nl, len(tag)):
                    own[loc[int(tag[g])]] += a[g]
                fx[key] = own if width > 1 else own[:, 0]
        if ms:
            fx["nspecial"] = d["nspecial"].reshape(nl, 3)
            fx["special"] = d["special"].reshape(nl, ms)
        if step == 0 and name in ("h2_default",):
            # full half list of the reference in canonical form: (i local idx, tag_j, shift code, special)
            ii = np.repeat(np.arange(nl, dtype=np.int32), d["numneigh"])
            jraw = d["neigh"]
            sb = (jraw >> 30) & 3
            j = jraw & 0x3FFFFFFF
            prd = d["boxhi"] - d["boxlo"]
            owner = np.array([loc[int(t)] for t in tag[j]], dtype=np.int32)
            shift = np.rint((x[j] - x[owner]) / prd).astype(np.int8)
            fx["half_i"] = ii.astype(np.int16)
            fx["half_j"] = owner.astype(np.int16)
            fx["half_shift"] = shift
            fx["half_special"] = sb.astype(np.int8)
            # tables of the reference (pins the table builder)
            for k in ("rtable", "drtable", "ftable", "dftable", "ctable", "dctable", "etable", "detable"):
                fx["tab_" + k] = first[k]
            fx["ncoulmask"] = int(first["ncoulmask"][0])
            fx["ncoulshiftbits"] = int(first["ncoulshiftbits"][0])
            fx["tabinnersq"] = first["tabinnersq"][0]
        np.savez_compressed(OUT / f"{name}_step{step}.npz", **fx)
        print(f"{name} step {step}: nlocal {nl} iterations {fx['iterations']} E_pol {fx['eng_pol']:.10g}")
    shutil.rmtree(work)
    return thermo


def parse_thermo(log):
    rows = []
    lines = log.splitlines()
    for i, l in enumerate(lines):
        if l.startswith("Step "):
            cols = l.split()
            for r in lines[i + 1:]:
                v = r.split()
                if len(v) != len(cols):
                    break
                try:
                    rows.append(dict(zip(cols, [float(t) for t in v])))
                except ValueError:
                    break
    return rows


def shipped_logs():
    """Thermo tables exactly as committed by the reference authors (strings, 8 significant digits)."""
    out = {}
    for key, rel in (("h2", "Bulk H2/log.lammps"), ("methane", "MOF5+Methane/log.lammps")):
        text = (EX / rel).read_text(errors="replace")
        lines = text.splitlines()
        for i, l in enumerate(lines):
            if l.startswith("Step "):
                cols = l.split()
                rows = []
                for r in lines[i + 1:]:
                    v = r.split()
                    if len(v) != len(cols):
                        break
                    try:
                        [float(t) for t in v]
                    except ValueError:
                        break
                    rows.append(dict(zip(cols, v)))
                out[key] = dict(source=f"polarization/examples/{rel}", columns=cols, rows=rows)
    return out


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    only = sys.argv[1:]
    if only == ["ewald"]:
        return ewald_goldens()
    if only == ["pppm"]:
        return ewald_goldens(kstyle="pppm")
    ours = {}
    for c in CASES:
        if only and c[0] not in only:
            continue
        ours[c[0]] = run_case(*c)
    if not only:
        logs = shipped_logs()
        logs["oracle_ref_runs"] = {k: v for k, v in ours.items()}
        (OUT / "thermo_logs.json").write_text(json.dumps(logs, indent=1))




# ---------------------------------------------------------------------------------------------------------------
# KSpace: reciprocal-space Ewald of the reference (src/KSPACE/ewald.cpp) as golden vectors.
# Forces and virial of KSpace alone = difference of two otherwise identical runs, with and without
# `kspace_modify compute no` (the pair forces of both runs are bit-identical).
# ---------------------------------------------------------------------------------------------------------------
NKTV2P_REAL = 68568.415  # src/update.cpp:161 (units real)


def ewald_case(name, data_text, cut_coul, accuracy, extra=(), kstyle="ewald"):
    work = Path(tempfile.mkdtemp(prefix=f"polgold_{name}_"))
    (work / "sys.data").write_text(data_text)
    outs = {}
    for tag, mod in (("on", ""), ("off", "kspace_modify compute no")):
        lines = ["units real", "boundary p p p", "atom_style full", "read_data sys.data", "mass * 1.0",
                 f"pair_style lj/cut/coul/long 2.5 {cut_coul}", "pair_coeff * * 0.0 1.0",
                 f"kspace_style {kstyle} {accuracy}", mod, *extra,
                 "thermo_style custom step elong ecoul pxx pyy pzz pxy pxz pyz vol",
                 "thermo_modify format float %.16g", f"dump d all custom 1 f_{tag}.dump id fx fy fz",
                 "dump_modify d format float %.17g sort id", "run 0"]
        (work / f"in.{tag}").write_text("\n".join(lines) + "\n")
        r = subprocess.run([str(LMP), "-in", f"in.{tag}", "-echo", "none", "-log", f"log.{tag}"], cwd=work,
                           capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout[-2000:])
            raise SystemExit(f"{name}: lmp_serial failed")
        log = (work / f"log.{tag}").read_text()
        th = parse_thermo(log)[0]
        rows = [l.split() for l in (work / f"f_{tag}.dump").read_text().splitlines()[9:]]
        f = np.array([[float(v) for v in r_[1:4]] for r_ in rows])
        outs[tag] = (th, f, log)
    th_on, f_on, log = outs["on"]
    th_off, f_off, _ = outs["off"]
    vol = th_on["Volume"]
    press = np.array([th_on[k] - th_off[k] for k in ("Pxx", "Pyy", "Pzz", "Pxy", "Pxz", "Pyz")])
    m = re.search(r"G vector \(1/distance\) = (\S+)", log)
    if kstyle == "pppm":
        gr = re.search(r"grid = (\d+) (\d+) (\d+)", log)
        so = re.search(r"stencil order = (\d+)", log)
        shutil.rmtree(work)
        return dict(elong=th_on["E_long"], f_kspace=f_on - f_off, virial_kspace=press * vol / NKTV2P_REAL,
                    g_ewald_printed=float(m.group(1)), grid=np.array([int(gr.group(i)) for i in (1, 2, 3)]),
                    order=int(so.group(1)), accuracy=accuracy, cut_coul=cut_coul)
    kv = re.search(r"KSpace vectors: actual max1d max3d = (\d+) (\d+) (\d+)", log)
    km = re.search(r"kxmax kymax kzmax\s+= (\d+) (\d+) (\d+)", log)
    shutil.rmtree(work)
    return dict(elong=th_on["E_long"], f_kspace=f_on - f_off, virial_kspace=press * vol / NKTV2P_REAL,
                g_ewald_printed=float(m.group(1)), kcount=int(kv.group(1)), kmax=int(kv.group(2)),
                kxyzmax=np.array([int(km.group(i)) for i in (1, 2, 3)]), accuracy=accuracy, cut_coul=cut_coul)


def ewald_goldens(kstyle="ewald"):
    """kstyle = "pppm": the same three systems through `kspace_style pppm` (src/KSPACE/pppm.cpp), plus an order-4 /
    explicit-mesh case; fixtures pppm_*.npz"""
    sys.path.insert(0, str(ROOT / "tests"))
    import polhelpers as H

    def data_of(x, q, typ, lo, hi):
        n = len(q)
        t = [f"ewald golden\n\n{n} atoms\n{int(typ.max())} atom types\n"]
        for d, c in enumerate("xyz"):
            t.append(f"{lo[d]:.17g} {hi[d]:.17g} {c}lo {c}hi")
        t.append("\nAtoms\n")
        t += [f"{i + 1} 0 {int(typ[i])} {q[i]:.17g} {x[i, 0]:.17g} {x[i, 1]:.17g} {x[i, 2]:.17g}" for i in range(n)]
        return "\n".join(t) + "\n"

    fx = H.load_fixture("h2_default_step0")
    cases = {"ewald_h2": (fx["x"], fx["q"], fx["type"], fx["boxlo"], fx["boxhi"], 10.797442, 1e-4),
             "ewald_methane": None, "ewald_brick": None}
    fm = H.load_fixture("methane_default_step0")
    cases["ewald_methane"] = (fm["x"], fm["q"], fm["type"], fm["boxlo"], fm["boxhi"], 12.8345, 1e-6)
    fl = H.lj_charge_fluid((6, 4, 3), seed=31)  # non-cubic box, 288 atoms
    cases["ewald_brick"] = (fl.x, fl.q, fl.type, fl.boxlo, fl.boxhi, 5.0, 1e-5)
    if kstyle == "pppm":
        cases = {k.replace("ewald_", "pppm_"): v + ((),) for k, v in cases.items()}
        cases["pppm_brick_order4"] = cases["pppm_brick"][:-1] + (("kspace_modify order 4 mesh 12 9 8 gewald 0.6",),)
        for name, (x, q, typ, lo, hi, cut, acc, extra) in cases.items():
            g = ewald_case(name, data_of(x, q, typ, lo, hi), cut, acc, extra=extra, kstyle="pppm")
            np.savez_compressed(OUT / f"{name}.npz", x=x, q=q, boxlo=lo, boxhi=hi, kspace_modify=np.array(" ".join(extra)), **g)
            print(f"{name}: n {len(q)} E_long {g['elong']:.12g} grid {g['grid']} order {g['order']} g {g['g_ewald_printed']}")
        return
    for name, (x, q, typ, lo, hi, cut, acc) in cases.items():
        g = ewald_case(name, data_of(x, q, typ, lo, hi), cut, acc)
        np.savez_compressed(OUT / f"{name}.npz", x=x, q=q, boxlo=lo, boxhi=hi, **g)
        print(f"{name}: n {len(q)} E_long {g['elong']:.12g} kcount {g['kcount']} kmax {g['kxyzmax']} g {g['g_ewald_printed']}")


if __name__ == "__main__":
    main()
