"""Static cross-checks between the MATLAB wrappers (subzero_b200/matlab/*.m) and the mex gateways they call.  MATLAB is not
installed, so the .m files cannot be executed; the gateways can (tests/test_zzzzz_mex_gateway.py).  What is left unverified is
the seam between the two -- a field the wrapper builds under one name and the gateway reads under another, a command that does
not exist, an output field that is never produced (round 1's recipe kept `isnan(x(i))` with `x` undefined,
floe_interactions_all.m:282).  These tests read both sides as text and hold the names together.  CPU only.
"""
import os
import re


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ML = os.path.join(ROOT, "subzero_b200", "matlab")


def _is_transpose(text, k):
    """MATLAB's apostrophe after an operand (identifier, closing bracket, another transpose) transposes; elsewhere it opens a string"""
    j = k - 1
    while j >= 0 and text[j] in " \t":
        j -= 1
    return j >= 0 and (text[j].isalnum() or text[j] in "_)]}.'") and text[k - 1] not in " \t,"


def _m(name):
    """a .m file with comments stripped and `...` continuations joined"""
    out = []
    for line in open(os.path.join(ML, name)).read().split("\n"):
        q, k = False, 0
        while k < len(line):                      # a % outside a '...' string starts a comment
            if line[k] == "'" and (q or not _is_transpose(line, k)):
                q = not q
            elif line[k] == "%" and not q:
                break
            k += 1
        out.append(line[:k].rstrip())
    return re.sub(r"\.\.\.\s*\n", " ", "\n".join(out))


def _struct_keys(text, var):
    """keys of `var = struct('k1', v1, 'k2', v2, ...)` (top-level arguments only)"""
    m = re.search(r"\b%s\s*=\s*struct\(" % re.escape(var), text)
    assert m, "no struct literal for %s" % var
    k, depth, args, cur, q = m.end(), 1, [], "", False
    while depth:
        c = text[k]
        if c == "'" and (q or not _is_transpose(text, k)):
            q = not q
        if not q:
            if c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
                if depth == 0:
                    break
            elif c == "," and depth == 1:
                args.append(cur.strip()); cur = ""; k += 1
                continue
        cur += c
        k += 1
    args.append(cur.strip())
    assert len(args) % 2 == 0, (var, args)
    keys = [a for a in args[0::2]]
    assert all(re.fullmatch(r"'\w+'", a) for a in keys), keys
    return [a.strip("'") for a in keys]


def _cpp(name):
    return open(os.path.join(ML, name)).read()


def _section(cpp, start, end=None):
    a = cpp.index(start)
    b = cpp.index(end, a + len(start)) if end else len(cpp)
    return cpp[a:b]


def _fields(section, var):
    """(required, optional) field names the gateway reads from struct argument `var`"""
    req = set(re.findall(r"need_field\(%s, \"(\w+)\"" % var, section)) | set(re.findall(r"scalar_field\(%s, \"(\w+)\", [^,]+, true\)" % var, section))
    opt = set(re.findall(r"opt_field\(%s, \"(\w+)\"" % var, section)) | set(re.findall(r"scalar_field\(%s, \"(\w+)\", [^,]+, false\)" % var, section))
    return req, opt - req


def _names(section):
    m = re.search(r"const char\* names\[\] = \{([^}]*)\}", section)
    return set(re.findall(r"\"(\w+)\"", m.group(1)))


def _uses(text, var):
    return set(re.findall(r"\b%s\.(\w+)" % var, text))


def test_contact_step_wrapper_and_gateway_agree_on_every_name():
    m, cpp = _m("sz_contact_step.m"), _cpp("sz_contact_mex.cpp")
    # the gateway binds prhs[0..2] to p, s, b
    for var, arg in (("prm", "p"), ("soa", "s"), ("bnd", "b")):
        keys = _struct_keys(m, var)
        req, opt = _fields(cpp, arg)
        assert len(set(keys)) == len(keys), keys
        assert req <= set(keys), (var, req - set(keys))
        assert set(keys) <= req | opt, (var, set(keys) - req - opt)
    assert set(_struct_keys(m, "soa")) == _fields(cpp, "s")[0]
    assert re.search(r"sz_contact_mex\(prm, soa\)", m) and re.search(r"sz_contact_mex\(prm, soa, bnd\)", m)
    produced = _names(cpp)
    assert _uses(m, "out") <= produced, _uses(m, "out") - produced
    # every per-floe output the reference's loop writes into Floe(i) is read from `out`
    assert {"rows", "row_off", "overlap_area", "fx", "fy", "torque", "alive", "xi", "yi", "kill", "transfer", "ghost_parent"} <= _uses(m, "out")
    # MATLAB is 1-based: row k of floe i is rows(:, row_off(i)+1 : row_off(i+1))
    assert "out.row_off(i)+1 : out.row_off(i+1)" in m and "out.row_off(N0+k)+1 : out.row_off(N0+k+1)" in m


def test_drop_in_uses_only_what_the_wrapper_returns_and_defines_what_the_tail_reads():
    m, w = _m("floe_interactions_all.m"), _m("sz_contact_step.m")
    ghost_fields = set(_struct_keys(w, "ghosts")) | set(re.findall(r"\bghosts\.(\w+)\s*=", w))
    assert _uses(m, "ghosts") <= ghost_fields, _uses(m, "ghosts") - ghost_fields
    # the reference's signature, argument for argument (floe_interactions_all.m:1)
    sig = re.search(r"function \[Floe,dissolvedNEW\] = floe_interactions_all\(([^)]*)\)", m).group(1)
    assert [a.strip() for a in sig.split(",")] == ["Floe", "floebound", "ocean", "winds", "c2_boundary", "dt", "HFo", "min_floe_size", "Nx", "Ny", "Nb", "dissolvedNEW",
                                                  "doInt", "COLLISION", "PERIODIC", "RIDGING", "RAFTING"]
    ref = "/root/reference/floe_interactions_all.m"
    if os.path.exists(ref):
        rsig = re.search(r"function \[Floe,dissolvedNEW\] = floe_interactions_all\(([^)]*)\)", open(ref).read()).group(1)
        assert [a.strip() for a in rsig.split(",")] == [a.strip() for a in sig.split(",")]
    # every variable handed to the tail, and the kill rule's x, is assigned (or an argument) before its use
    body = m[m.index(")", m.index("function")) + 1:]
    defined = set(a.strip() for a in sig.split(","))
    for line in body.split("\n"):
        lhs = re.match(r"\s*(?:\[([^\]]+)\]|(\w+))\s*=[^=]", line)
        call = re.search(r"sz_floe_interactions_tail\(([^)]*)\)", line)
        if call:
            for a in call.group(1).split(","):
                assert a.strip() in defined, "the tail is handed %r, which is not defined at that point" % a.strip()
        if "isnan(x(i))" in line:
            assert "x" in defined
        if lhs:
            for v in (lhs.group(1) or lhs.group(2)).replace("~", " ").split(","):
                defined.add(v.strip().split("(")[0].split(".")[0])
    assert "x = cat(1,Floe.Xi);" in m and m.index("x = cat(1,Floe.Xi);") < m.index("sz_contact_step(")        # the PRE-step centroid (:71,282)


def test_resident_wrapper_and_gateway_agree_on_every_name():
    m, cpp = _m("sz_resident_timestep.m"), _cpp("sz_resident_mex.cpp")
    commands = set(re.findall(r"cmd == \"(\w+)\"", cpp))
    used = set(re.findall(r"sz_resident_mex\('(\w+)'", m))
    assert used and used <= commands, used - commands
    assert {"upload", "step", "floe_outputs", "rows", "trajectory_init", "set_ocean", "set_points", "ocean_forcing", "trajectory_step", "state"} <= used
    # every command the gateway's header comment documents exists, and the other way round
    documented = set(re.findall(r"sz_resident_mex\('(\w+)'", cpp[:cpp.index("#include")]))
    assert documented <= commands and commands - documented <= {"weld_search", "simplify_search"}, (documented ^ commands)
    up = _section(cpp, "static void cmd_upload", "void mexFunction")
    for var, arg in (("prm", "p"), ("soa", "s"), ("bnd", "b")):
        keys, (req, opt) = set(_struct_keys(m, var)), _fields(up, arg)
        assert req <= keys <= req | opt, (var, req - keys, keys - req - opt)
    ti = _section(cpp, 'if (cmd == "trajectory_init")', 'if (cmd == "set_ocean")')
    keys, (req, opt) = set(_struct_keys(m, "st")), _fields(ti, "s")
    assert req <= keys <= req | opt, (req - keys, keys - req - opt)
    keys, (req, opt) = set(_struct_keys(m, "tp")), _fields(_section(cpp, "static SzTrajectoryParams traj_params", "static void cmd_upload"), "tp")
    assert req <= keys <= req | opt, (req - keys, keys - req - opt)
    assert _uses(m, "out") <= _names(_section(cpp, 'if (cmd == "floe_outputs")', 'if (cmd == "rows")'))
    assert _uses(m, "r") <= _names(_section(cpp, 'if (cmd == "rows")', 'if (cmd == "trajectory_init")'))
    state = _names(_section(cpp, 'if (cmd == "state")', 'if (cmd == "fracture_deform")'))
    assert _uses(m, "s") <= state, _uses(m, "s") - state
    # everything the integrator advances goes back into Floe(i) (calc_trajectory.m:170-234)
    assert {"x", "y", "u", "v", "ksi", "h", "alive", "mass", "inertia", "alpha", "dXi_p", "dYi_p", "dUi_p", "dVi_p", "dalpha_p", "dksi_p", "stress", "strain",
            "FxOA", "FyOA", "torqueOA", "cax", "cay", "flags"} <= _uses(m, "s")
    # set_ocean reads the reference's own struct fields (Initialize_Model/initialize_ocean.m:4-33, Subzero.m:46-49)
    so = _section(cpp, 'if (cmd == "set_ocean")', 'if (cmd == "set_points")')
    assert _fields(so, "o")[0] == {"Xo", "Yo", "Uocn", "Vocn", "fCoriolis", "turn_angle"} and _fields(so, "w")[0] == {"u", "v"}
    ref = "/root/reference/Initialize_Model/initialize_ocean.m"
    if os.path.exists(ref):
        have = set(re.findall(r"\bocean\.(\w+)\s*=", open(ref).read()))
        assert _fields(so, "o")[0] <= have
        assert {"u", "v"} <= set(re.findall(r"\bwinds\.(\w+)\s*=", open("/root/reference/Subzero.m").read()))


def test_python_mirror_builds_the_same_structs_as_the_wrapper():
    """tests/mexmock.py (what the executed gateway tests feed mexFunction) and sz_contact_step.m build prm / soa / bnd with the
    same field names, so the executed test exercises exactly the wrapper's calling convention"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mexmock
    import subzero_b200 as sz
    prm, soa = sz.voronoi_field(20, seed=1)
    m = _m("sz_contact_step.m")
    assert set(mexmock.prm_struct(prm)) == set(_struct_keys(m, "prm"))
    assert set(mexmock.soa_struct(soa)) == set(_struct_keys(m, "soa"))
    b = sz.abi.Boundary([0, 1, 1], [0, 0, 1], [0, 1, 1], [0, 0, 1], 1.0)
    assert set(mexmock.bnd_struct(b)) == set(_struct_keys(m, "bnd"))
