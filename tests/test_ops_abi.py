"""The ctypes prototypes in ``mdi_llm_b200/ops/__init__.py`` against the ``extern "C"`` signatures in ``csrc/*.cu``:
parameter count and kind of every entry point.  (A mismatch is silent memory corruption at run time: ctypes pushes what
the prototype says, the kernel launcher reads what its C signature says.)  Runs on CPU — the library only has to load."""
import ctypes
import re
from pathlib import Path

import pytest

CSRC = Path(__file__).resolve().parents[1] / "mdi_llm_b200" / "ops" / "csrc"


def _c_signatures():
    sigs = {}
    for f in sorted(CSRC.glob("*.cu")):
        src = re.sub(r"//[^\n]*", "", f.read_text())
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        # every top-level definition `<type> mdi_xxx(<params>) {` — the exported entry points all carry the prefix
        for m in re.finditer(r'^(?:extern\s+"C"\s+)?(?:[A-Za-z_][\w \t\*]*?)\b(mdi_\w+)\s*\(([^{;]*?)\)\s*\{', src, flags=re.M | re.S):
            name, params = m.group(1), m.group(2).strip()
            sigs[name] = [] if params in ("", "void") else [" ".join(p.split()) for p in params.split(",")]
    return sigs


def _kind(c_param: str) -> str:
    if "*" in c_param or re.search(r"\bcuda\w+_t\b", c_param):  # CUDA handle types (stream, graph exec, ...) are pointers
        return "ptr"
    t = c_param.rsplit(" ", 1)[0] if " " in c_param else c_param
    t = t.replace("const ", "").strip()
    return {"int": "i32", "unsigned int": "u32", "unsigned": "u32", "long long": "i64", "unsigned long long": "u64", "size_t": "u64",
            "float": "f32", "double": "f64"}.get(t, f"?{t}")


_CT = {ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr", ctypes.c_int: "i32", ctypes.c_uint: "u32", ctypes.c_longlong: "i64",
       ctypes.c_int64: "i64", ctypes.c_ulonglong: "u64", ctypes.c_uint64: "u64", ctypes.c_size_t: "u64", ctypes.c_float: "f32",
       ctypes.c_double: "f64"}


def test_every_prototype_matches_its_c_signature():
    from mdi_llm_b200 import ops

    try:
        lib = ops.lib()
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"kernel library not loadable here: {e}")
    sigs = _c_signatures()
    assert len(sigs) > 40, sorted(sigs)
    checked, problems = 0, []
    for name, params in sorted(sigs.items()):
        fn = getattr(lib, name)
        proto = fn.argtypes
        if proto is None:
            if params:  # an entry point called without a declared prototype relies on ctypes' int default: only fine without arguments
                problems.append(f"{name}: {len(params)} parameters but no argtypes declared")
            continue
        want = [_kind(p) for p in params]
        got = []
        for t in proto:
            k = _CT.get(t)
            if k is None and isinstance(t, type) and issubclass(t, ctypes._Pointer):
                k = "ptr"
            got.append(k or f"?{t}")
        if len(want) != len(got):
            problems.append(f"{name}: C has {len(want)} parameters, prototype has {len(got)}")
        else:
            bad = [(i, params[i], w, g) for i, (w, g) in enumerate(zip(want, got)) if w != g]
            if bad:
                problems.append(f"{name}: " + "; ".join(f"#{i} `{p}` is {w}, prototype says {g}" for i, p, w, g in bad))
        checked += 1
    assert not problems, "\n".join(problems)
    assert checked > 35


def test_the_check_notices_a_drifted_prototype():
    from mdi_llm_b200 import ops

    try:
        lib = ops.lib()
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"kernel library not loadable here: {e}")
    saved = lib.mdi_qkv_decode.argtypes
    try:
        lib.mdi_qkv_decode.argtypes = list(saved)[:-1]  # "someone added a parameter on the C side only"
        with pytest.raises(AssertionError, match="mdi_qkv_decode"):
            test_every_prototype_matches_its_c_signature()
        wrong = list(saved)
        wrong[9] = ctypes.c_int  # x_slot_stride is a long long
        lib.mdi_qkv_decode.argtypes = wrong
        with pytest.raises(AssertionError, match="x_slot_stride"):
            test_every_prototype_matches_its_c_signature()
    finally:
        lib.mdi_qkv_decode.argtypes = saved


def test_constants_mirrored_in_python_match_the_headers():
    """Step-descriptor layout, poison value and activation codes exist on both sides of the ctypes boundary."""
    from mdi_llm_b200 import ops

    h = (CSRC / "common.cuh").read_text()
    defs = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"^#define\s+(MDI_\w+)\s+(0x[0-9a-fA-F]+|\d+)\b", h, flags=re.M)}
    for py, c in (("CTX_SLOT", "MDI_CTX_SLOT"), ("CTX_POS", "MDI_CTX_POS"), ("CTX_WAIT", "MDI_CTX_WAIT"), ("CTX_SIGNAL", "MDI_CTX_SIGNAL"),
                  ("CTX_TOKEN", "MDI_CTX_TOKEN"), ("CTX_STEP", "MDI_CTX_STEP"), ("CTX_INTS", "MDI_CTX_INTS"), ("POISON", "MDI_POISON")):
        assert getattr(ops, py) == defs[c], (py, c)
    enum = re.search(r"enum Act \{([^}]*)\}", h).group(1)
    codes = {k.strip()[4:].lower(): int(v) for k, v in (item.split("=") for item in enum.split(","))}
    assert codes == ops.ACT
    from mdi_llm_b200.parallel.protocol_model import POISON

    assert POISON == ops.POISON


# (entry point, C parameter) -> a token that must appear in the Python argument expression instead of the C name
_ALIASES = {("mdi_set_prefill_attn_pipe", "on"): "mode", ("mdi_linear_decode", "pf_a"): "prefetch", ("mdi_linear_decode", "pf_b"): "prefetch",
            ("mdi_linear_decode", "pf_bytes"): "prefetch", ("mdi_qkv_decode", "k"): "shape", ("mdi_embed", "c"): "shape",
            ("mdi_rmsnorm_rows", "rows"): "numel", ("mdi_sample", "v"): "vocab", ("mdi_sample_fast", "v"): "vocab",
            ("mdi_quantize_rows_fp8", "ld_s"): "m_pad", ("mdi_gemm_fp8", "ld_as"): "a_scale_t", ("mdi_gemm_fp8", "c"): "c_ptr",
            ("mdi_attn_prefill", "q_scratch"): "q_s", ("mdi_attn_prefill", "max_seq"): "s", ("mdi_graph_end", "n_nodes"): "n"}


def test_wrappers_pass_their_arguments_in_the_c_order():
    """Prototypes only fix count and kind; two `long long` strides swapped at a call site would still corrupt silently.
    Every positional argument of every `lib().mdi_*()` call in ops/__init__.py must mention its C parameter's name (or a
    listed alias): an argument in the wrong position names the wrong parameter."""
    import ast

    sigs = _c_signatures()
    src = (CSRC.parent / "__init__.py").read_text()

    def tokens(node):
        out = set()
        for n in ast.walk(node):
            if isinstance(n, ast.Name):
                out.add(n.id.lower())
            elif isinstance(n, ast.Attribute):
                out.add(n.attr.lower())
        return out

    checked, problems = 0, []
    for n in ast.walk(ast.parse(src)):
        if not (isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute) and n.func.attr in sigs):
            continue
        name, params = n.func.attr, sigs[n.func.attr]
        if len(params) < 2:
            continue  # nothing to get out of order
        pairs = list(zip(n.args, params))
        star = [i for i, a in enumerate(n.args) if isinstance(a, ast.Starred)]
        if star:  # one tuple of tuning knobs forwarded as a block: what precedes it aligns from the front, the rest from the back
            assert len(star) == 1
            i, tail = star[0], len(n.args) - star[0] - 1
            pairs = list(zip(n.args[:i], params[:i])) + list(zip(n.args[i + 1:], params[len(params) - tail:]))
        elif len(n.args) != len(params):
            problems.append(f"{name}: call passes {len(n.args)} arguments, C takes {len(params)}")
            continue
        for a, p in pairs:
            pname = p.split()[-1].strip("*").lower()
            flat = " ".join(sorted(tokens(a)))
            want = _ALIASES.get((name, pname))
            ok = (want in flat) if want else (pname in flat or any(len(x) > 1 and x in flat for x in pname.split("_")))
            checked += 1
            if not ok:
                problems.append(f"{name}: parameter `{pname}` receives `{ast.unparse(a)[:60]}`")
    assert not problems, "\n".join(problems)
    assert checked > 250


def test_built_library_contains_the_blackwell_instructions():
    """The shipped binary really is the sm_100a build: tcgen05 MMAs (UTCHMMA bf16, UTCQMMA fp8), TMEM loads (LDTM), TMA tensor
    loads (UTMALDG) and bulk copies (UBLKCP), mbarrier ops (SYNCS), cluster barriers — and no legacy tensor-core path (HMMA)
    in the GEMM / attention kernels."""
    import shutil
    import subprocess

    from mdi_llm_b200.ops import build

    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(tool).exists() or not build.LIB.exists():
        pytest.skip("cuobjdump or the built library is not available here")
    out = subprocess.run([tool, "-sass", str(build.LIB)], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in out
    per_fn, fn = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            per_fn[fn] = set()
        elif fn:
            m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
            if m:
                per_fn[fn].add(m.group(1).split(".")[0])

    def ops_of(substr):
        hit = [v for k, v in per_fn.items() if substr in k]
        assert hit, substr
        return set().union(*hit)

    gemm, fp8, attn = ops_of("gemm_bf16_tcgen05"), ops_of("gemm_fp8_blockscaled"), ops_of("attn_prefill_tcgen05")
    assert {"UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "SYNCS"} <= gemm
    assert {"UTCQMMA", "LDTM", "UTMALDG", "SYNCS"} <= fp8
    assert {"UTCHMMA", "LDTM", "UTMALDG"} <= attn
    assert not ({"HMMA", "IMMA", "QMMA"} & (gemm | fp8 | attn))  # no mma.sync-class fallback
    assert {"UBLKCP", "SYNCS"} <= ops_of("stream_bulk_kernel")   # decode weight streaming through the TMA engine
    assert "UCGABAR_ARV" in ops_of("attn_decode_kernel") or "UCGABAR_WAIT" in ops_of("attn_decode_kernel")  # cluster barrier (DSMEM merge)


def test_hot_kernels_fit_their_occupancy_budgets():
    """Register / stack budgets the launch configurations rely on: the decode streamers run 3 CTAs of 256 threads per SM
    (<= 85 registers, no stack), the flagship attention instantiation and the MoE passes do not spill."""
    import shutil
    import subprocess

    from mdi_llm_b200.ops import build

    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(tool).exists() or not build.LIB.exists():
        pytest.skip("cuobjdump or the built library is not available here")
    out = subprocess.run([tool, "-res-usage", str(build.LIB)], capture_output=True, text=True, timeout=300).stdout
    usage = {m.group(1): {k: int(v) for k, v in re.findall(r"(REG|STACK|LOCAL):(\d+)", m.group(2))}
             for m in re.finditer(r"Function (\S+):\s*\n\s*(REG:[^\n]*)", out)}
    assert len(usage) > 60

    def sel(substr):
        hit = {k: v for k, v in usage.items() if substr in k}
        assert hit, substr
        return hit

    for k, u in {**sel("stream_bulk_kernelILi0ELi2"), **sel("stream_bulk_kernelILi1ELi2"), **sel("stream_bulk_kernelILi2ELi2"),
                 **sel("moe_stream_kernel"), **sel("moe_bulk_kernel")}.items():
        assert u["REG"] * 256 * 3 <= 65536 and u["STACK"] == 0 and u["LOCAL"] == 0, (k, u)
    for k, u in sel("attn_decode_kernelILi128ELi4").items():  # Llama-3 / Mistral: 4 query heads per KV head
        assert u["STACK"] == 0 and u["LOCAL"] == 0, (k, u)
    for k, u in {**sel("gemm_bf16_tcgen05"), **sel("gemm_fp8_blockscaled")}.items():
        assert u["LOCAL"] == 0 and u["REG"] * 384 <= 65536, (k, u)  # one CTA per SM of up to 384 threads (fp8: 8 promotion warps)
