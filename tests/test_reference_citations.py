"""CPU, only where /root/reference exists (this container; skipped on the GPU box): the reference statements the oracle
restates, at the file:line the oracle cites, still read the way the restatement assumes -- the comparison operators of the
accept tests (`>` return, `>=` in the batched clock), the proposal roundings (`floor`, `ceiling`), the neighbour sets, the halo
copies, the observable sums -- and oracle/oracle.c uses the same operators.  The reference has no tests or golden vectors and
cannot be built here (no Fortran compiler): this ties the restatement to the reference's TEXT, the one artefact available.
Whitespace is ignored in the comparison."""
import os
import re

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="the reference checkout is not present")


def _norm(s):
    return re.sub(r"\s+", "", s.replace("&", ""))


def _lines(path, lo, hi):
    with open(os.path.join(REF, path)) as f:
        src = f.read().splitlines()
    return _norm("".join(src[lo - 1:hi]))


# (file, first line, last line, statements that must appear in that range)
CITED = [
    # Ising 2D: table, delta energy, accept test, flip, halo, random start, observables
    ("src/ising2d_gpu_m.f90", 126, 130, ["this%exparr_(:)=1.0_real64", "this%exparr_(diff)=exp(-this%beta()*diff)"]),
    ("src/ising2d_gpu_m.f90", 191, 196, ["res=2*spins(idx)*(spins(idx+1)+spins(idx-1)+spins(idx+nx)+spins(idx-nx))"]),
    ("src/ising2d_gpu_m.f90", 155, 161, ["idx=2*((blockIdx%x-1)*blockDim%x+threadIdx%x)-2+offset", "if(randoms(idx)>exparr(delta_energy))return",
                                         "spins(idx)=-spins(idx)"]),
    ("src/ising2d_gpu_m.f90", 100, 106, ["spins(nall+idx)=spins(idx)", "spins(idx-nx)=spins(nall-nx+idx)"]),
    ("src/ising2d_gpu_m.f90", 83, 83, ["spins(idx)=merge(1,-1,randoms(idx)<0.5_real64)"]),
    ("src/ising2d_gpu_m.f90", 205, 228, ["res=res-int(spins(i)*(spins(i+1)+spins(i+this%nx_)),int64)", "res=res+spins(i)"]),
    # Ising 3D
    ("src/ising3d_gpu_m.f90", 196, 205, ["spins(idx-1)+spins(idx+1)+spins(idx-nx)+spins(idx+nx)+spins(idx-nxy)+spins(idx+nxy)",
                                         "if(randoms(idx)>ws(sum_spin,spins(idx)))return", "spins(idx)=1-spins(idx)"]),
    ("src/ising3d_gpu_m.f90", 153, 171, ["e1=this%energy_table_(s1,0)+this%energy_table_(s2,0)", "this%ws_(s1+s2,0)=min(1.0_real64,exp(-this%beta_*(e2-e1)))",
                                         "this%ws_(s1+s2,1)=min(1.0_real64,exp(-this%beta_*(e1-e2)))"]),
    ("src/ising3d_gpu_m.f90", 111, 122, ["spins(nall+idx)=spins(idx)", "spins(idx-nxy)=spins(nall-nxy+idx)"]),
    ("src/ising3d_gpu_m.f90", 99, 99, ["spins(idx)=merge(1,0,randoms(idx)<0.5_real64)"]),
    ("src/ising3d_gpu_m.f90", 245, 276, ["res=res+energy_table(spins(i+1)+spins(i+this%nx_)+spins(i+this%nxy_),spins(i))", "res=res+spins(i)"]),
    # clock, helical: proposal floor(u q), accept `>` return (single) and `>=` return (batched)
    ("src/clock_gpu_m.f90", 211, 215, ["next_state=floor(next_states(idx)*max_state)",
                                       "if(randoms(idx)>ws(spins(idx+nx),spins(idx-nx),spins(idx-1),spins(idx+1),spins(idx),next_state))return",
                                       "spins(idx)=next_state"]),
    ("src/clock_gpu_multi_m.f90", 230, 235, ["next_state=floor(next_states(idx_x,idx_y)*max_state)", "if(randoms(idx_x,idx_y)>=ws("]),
    ("src/clock_gpu_m.f90", 103, 103, ["spins(idx)=floor(randoms(idx)*max_state)"]),
    # clock, periodic tableall / dual lattice: new = c + ceiling(u1 (q - 1)), accept iff u2 <= prob
    ("src/clock/clock_tableall_gpu_m.f90", 140, 150, ["new_state=sixclock(x,y)+ceiling(rnds(1,x,y)*(mstate-1))", "if(rnds(2,x,y)<=prob)"]),
    ("src/clock/clock_dual_lattice_tableall_m.f90", 142, 153, ["ceiling(rnds(1,actual_x,y)*(mstate-1))", "if(rnds(2,actual_x,y)<=prob)"]),
    # XY periodic: candidate, delta energy, accept, over-relaxation
    ("src/xy2d_periodic_gpu_m.f90", 382, 385, ["candidate(1:2)=[cos(2*pi*candidates(x,y)),sin(2*pi*candidates(x,y))]",
                                               "if(randoms(x,y)>exp(-beta*delta_energy))return"]),
    ("src/xy2d_periodic_gpu_m.f90", 390, 397, ["center_diff=candidate(:)-spins(x,y,:)", "neighbor_summ=spins(x+1,y,:)+spins(x-1,y,:)+spins(x,y+1,:)+spins(x,y-1,:)",
                                               "res=-(center_diff(1)*neighbor_summ(1)+center_diff(2)*neighbor_summ(2))"]),
    ("src/xy2d_periodic_gpu_m.f90", 426, 438, ["abs_local_field_inv=1/hypot(local_field(1),local_field(2))",
                                               "spins(x,y,1:2)=(2*sum(local_field(1:2)*spins(x,y,1:2)))*local_field(1:2)-spins(x,y,1:2)",
                                               "spins(x,y,1:2)=spins(x,y,1:2)/rabs"]),
    # clock tables (loop nest order and the floating-point expression order are what the oracle copies)
    ("src/clock_gpu_m.f90", 105, 146, ["docenter_after=0,this%max_state_-1docenter_before=0,this%max_state_-1dol=0,this%max_state_-1dok=0,this%max_state_-1doj=0,this%max_state_-1doi=0,this%max_state_-1",
                                       "delta_e=(energy_table(i,j,center_after)+energy_table(k,l,center_after))-(energy_table(i,j,center_before)+energy_table(k,l,center_before))",
                                       "if(delta_e<=0.0_real64)then", "ws(i,j,k,l,center_before,center_after)=exp(-this%beta_*delta_e)",
                                       "res=-sum(cos(this%pi_state_inv_*[i-center,j-center]))"]),
    ("src/clock/clock_tableall_gpu_m.f90", 66, 86, ["delta_e=state_center_right_up_to_energy(new_c,r,u)-state_center_right_up_to_energy(c,r,u)+state_center_right_up_to_energy(new_c,l,d)-state_center_right_up_to_energy(c,l,d)",
                                                    "if(delta_e<=0d0)then", "probability_h(c,new_c,r,u,l,d)=exp(-beta*delta_e)"]),
    # clock and XY observables
    ("src/clock_gpu_m.f90", 245, 280, ["res=res+energy_table(spins(i-this%nx_),spins(i-1),spins(i))", "res=res+spin_magne(spins(i))"]),
    ("src/clock/clock_tableall_gpu_m.f90", 155, 181, ["res=res+state_to_magne(sixclock(x,y))", "res=res*nall_inv",
                                                      "res=res+state_center_right_up_to_energy(sixclock(x,y),sixclock(rx,y),sixclock(x,uy))"]),
    ("src/xy2d_periodic_gpu_m.f90", 496, 534, ["res=res-spins(x,y,1)*(spins(x+1,y,1)+spins(x,y+1,1))", "res=res-spins(x,y,2)*(spins(x+1,y,2)+spins(x,y+1,2))",
                                               "res=res+spins(x,y,1)", "res=res+spins(x,y,2)"]),
    ("src/xy2d_periodic_gpu_m.f90", 121, 121, ["spins(x,y,:)=[cos(2*pi*randoms(x,y)),sin(2*pi*randoms(x,y))]"]),
]


@pytest.mark.parametrize("path,lo,hi,stmts", CITED, ids=[f"{c[0].split('/')[-1]}:{c[1]}" for c in CITED])
def test_cited_reference_lines_read_as_restated(path, lo, hi, stmts):
    text = _lines(path, lo, hi)
    for s in stmts:
        assert _norm(s) in text, (path, lo, hi, s)


def test_oracle_uses_the_reference_operators():
    """the same statements in oracle/oracle.c: `>` skips (accept iff u <= w), `>=` in the batched clock, floor / ceil proposals"""
    c = _norm(open(os.path.join(ROOT, "oracle", "oracle.c")).read())
    for s in ["if(randoms[idx-1]>ws[sum_spin+7*I3(idx)])continue;",           # src/ising3d_gpu_m.f90:203
              "I3(idx)=1-I3(idx);",                                             # :205
              "I3(nall+idx)=I3(idx);", "I3(idx-nxy)=I3(nall-nxy+idx);",         # :119-121
              "(randoms[idx-1]<0.5)?1:0",                                       # :99
              "(randoms[idx-1]<0.5)?1:-1"]:                                     # src/ising2d_gpu_m.f90:83
        assert _norm(s) in c, s
    # table builders: the floating-point expression order of the reference (src/clock_gpu_m.f90:128-129,
    # src/clock/clock_tableall_gpu_m.f90:72-75)
    assert _norm("double de = (ET(i, j, ca) + ET(k, l, ca)) - (ET(i, j, cb) + ET(k, l, cb));") in c
    assert _norm("double de = E3(n, r, u) - E3(c, r, u) + E3(n, l, d) - E3(c, l, d);") in c
    assert re.search(r"floor\(", c) and re.search(r"ceil\(", c)
    # the batched clock's strict comparison is a separate code path (`>=` return <=> accept iff u < w)
    assert ">=" in c


# ---- the drop-in boundary: module, type and procedure names of the shims against the reference's own modules ----
SHIMS = os.path.join(ROOT, "cuda_fortran_mc_simulation_spin_b200", "fortran")
TYPED = [("src/ising2d_gpu_m.f90", "ising2d_gpu_m.f90"), ("src/ising3d_gpu_m.f90", "ising3d_gpu_m.f90"),
         ("src/clock_gpu_m.f90", "clock_gpu_m.f90"), ("src/clock_gpu_multi_m.f90", "clock_gpu_multi_m.f90"),
         ("src/xy2d_periodic_gpu_m.f90", "xy2d_periodic_gpu_m.f90"), ("src/xy2d_gpu_m.f90", "xy2d_gpu_m.f90")]
PROCEDURAL = [("src/clock/clock_tableall_gpu_m.f90", "clock/clock_tableall_gpu_m.f90"),
              ("src/clock/clock_dual_lattice_tableall_m.f90", "clock/clock_dual_lattice_tableall_m.f90"),
              ("src/clock/clock_table_gpu_m.f90", "clock/clock_table_gpu_m.f90"),
              ("src/clock/clock_simple_gpu_m.f90", "clock/clock_simple_gpu_m.f90")]


def _interface(src):
    """(module name, public type names, public type-bound procedure names, names on `public ::` lines)"""
    mod = re.search(r"^\s*module\s+(\w+)", src, flags=re.M | re.I).group(1).lower()
    types = {m.lower() for m in re.findall(r"^\s*type\s*(?:,\s*public\s*)?::\s*(\w+)", src, flags=re.M | re.I)}
    bound = set()
    for m in re.finditer(r"^\s*procedure\s*,\s*pass\s*(,\s*private\s*)?::\s*(\w+)", src, flags=re.M | re.I):
        if not m.group(1):
            bound.add(m.group(2).lower())
    public = set()
    for m in re.finditer(r"^\s*(?:[\w()=, ]*,\s*)?public\s*(?:,\s*protected\s*)?::\s*(.+)$", src, flags=re.M | re.I):
        for name in re.split(r",", re.sub(r"=[^,]*", "", m.group(1))):
            name = name.strip().split("!")[0].strip()
            if re.fullmatch(r"\w+", name):
                public.add(name.lower())
    return mod, types, bound, public


@pytest.mark.parametrize("ref,shim", TYPED + PROCEDURAL, ids=[s for _, s in TYPED + PROCEDURAL])
def test_shim_keeps_the_reference_interface(ref, shim):
    """same module name; every public type, every public type-bound procedure and every name the reference module makes public
    exists under the same name in the shim (the shims add procedures, they never rename or drop one)"""
    rmod, rtypes, rbound, rpublic = _interface(open(os.path.join(REF, ref)).read())
    smod, stypes, sbound, spublic = _interface(open(os.path.join(SHIMS, shim)).read())
    assert smod == rmod
    assert rtypes & rpublic <= stypes, (rtypes & rpublic) - stypes
    assert rbound <= sbound, sorted(rbound - sbound)
    assert rpublic <= spublic | stypes, sorted(rpublic - spublic - stypes)


def _signatures(src):
    """public type-bound procedure -> (subroutine | function, result type, dummy argument names, {argument: (type + attributes, shape)})"""
    src = re.sub(r"&\s*\n\s*&?", "", src)
    bound = {m.group(2).lower(): m.group(3).lower()
             for m in re.finditer(r"^\s*procedure\s*,\s*pass\s*(,\s*private\s*)?::\s*(\w+)\s*=>\s*(\w+)", src, flags=re.M | re.I) if not m.group(1)}
    out = {}
    for name, impl in bound.items():
        m = re.search(r"^[^\n!]*(subroutine|function)\s+" + impl + r"\s*\(([^)]*)\)[^\n]*\n(.*?)^\s*end\s+(subroutine|function)", src,
                      flags=re.I | re.S | re.M)
        if not m:
            continue
        head = m.group(0).split("\n")[0].lower()
        args = [a.strip().lower() for a in m.group(2).split(",")]
        decl = {}
        for line in m.group(3).split("\n"):
            mm = re.match(r"\s*([^!:]+?)\s*::\s*([^!]+)", line)
            if not mm:
                continue
            typ = re.sub(r"\s+", "", mm.group(1).lower())
            for v in re.split(r",(?![^()]*\))", mm.group(2)):
                v = re.sub(r"\s+", "", v.lower())
                b = re.match(r"\w+", v)
                if b and b.group(0) in args and b.group(0) not in decl:     # the first declaration: contained procedures reuse names
                    decl[b.group(0)] = (typ, v[len(b.group(0)):])
        res = re.search(r"(integer\(\w+\)|real\(\w+\)|logical)\s+function", head)
        out[name] = (m.group(1).lower(), res.group(1) if res else None, args, decl)
    return out


@pytest.mark.parametrize("ref,shim", TYPED, ids=[s for _, s in TYPED])
def test_shim_procedures_keep_the_reference_signatures(ref, shim):
    """every public type-bound procedure of the reference: same kind (subroutine / function), same result type, same dummy
    argument names in the same order, same declared type, kind, intent and shape for every argument but the passed object"""
    a = _signatures(open(os.path.join(REF, ref)).read())
    b = _signatures(open(os.path.join(SHIMS, shim)).read())
    assert len(a) >= 10
    for name, (kind, res, args, decl) in a.items():
        assert name in b, name
        kind2, res2, args2, decl2 = b[name]
        assert (kind, res, args) == (kind2, res2, args2), (name, (kind, res, args), (kind2, res2, args2))
        for arg, t in decl.items():
            if arg != "this":
                assert decl2.get(arg) == t, (name, arg, t, decl2.get(arg))
