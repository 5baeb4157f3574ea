"""rust/woxel-b200-sys (uncompiled here: no rustc in the image) against include/woxel_b200.h: every #[repr(C)] struct has the
header's fields in the header's order with matching types, every function of the header is declared in the extern "C" block
with matching argument and return types, and the status / option constants agree.  Mirrors what the reference's own types
fix: ComputeState's field order (src/render/gpu_types/compute_state.rs:9-29)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "include", "woxel_b200.h")).read()
RS = open(os.path.join(ROOT, "rust", "woxel-b200-sys", "src", "lib.rs")).read()

C_SCALAR = {"uint8_t": "u8", "int32_t": "i32", "uint32_t": "u32", "uint64_t": "u64", "int64_t": "i64", "float": "f32", "int": "c_int",
            "size_t": "usize", "char": "c_char", "void": "c_void"}


def strip_comments(s):
    return re.sub(r"/\*.*?\*/", "", s, flags=re.S)


def c_type_to_rust(ctype: str, array: str | None = None, as_param: bool = False) -> str:
    ctype = ctype.strip()
    const = "const" in ctype.split()
    base = " ".join(t for t in ctype.replace("*", " ").split() if t != "const")
    stars = ctype.count("*")
    rust = C_SCALAR.get(base, base)  # struct names map to themselves
    for k in range(stars):
        # only the innermost pointee carries `const` in this header (const T *p, T **out)
        rust = ("*const " if (const and k == 0) else "*mut ") + rust
    if array:
        rust = f"*const {rust}" if (as_param and const) else (f"*mut {rust}" if as_param else f"[{rust}; {array}]")
    return rust


def c_structs():
    out = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} \1;", strip_comments(HDR), flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            # "const uint64_t *kids5" / "uint32_t n5, n4, n3" / "float view_proj[16]"
            mm = re.match(r"(.*?)([\w\[\], \*]+)$", decl)
            head = decl
            names = []
            first = re.match(r"((?:const\s+)?\w+)\s*(.*)$", head)
            ctype, rest = first.group(1), first.group(2)
            for part in rest.split(","):
                part = part.strip()
                stars = part.count("*")
                name = part.replace("*", "").strip()
                arr = None
                am = re.match(r"(\w+)\[(\d+)\]$", name)
                if am:
                    name, arr = am.group(1), am.group(2)
                fields.append((name, c_type_to_rust(ctype + " " + "*" * stars, arr)))
        out[m.group(1)] = fields
    return out


def rust_structs():
    out = {}
    for m in re.finditer(r"#\[repr\(C\)\]\s*(?:#\[derive\([^\]]*\)\]\s*)?pub struct (\w+) \{(.*?)\n\}", RS, flags=re.S):
        fields = []
        for line in m.group(2).split("\n"):
            line = line.split("//")[0].strip().rstrip(",")
            fm = re.match(r"(?:pub )?(\w+): (.+)$", line)
            if fm:
                fields.append((fm.group(1), fm.group(2).strip()))
        out[m.group(1)] = fields
    return out


def c_functions():
    out = {}
    body = strip_comments(HDR)
    body = re.sub(r"typedef (struct|enum) \w+ \{.*?\} \w+;", "", body, flags=re.S)
    for m in re.finditer(r"^((?:const\s+)?\w+\s*\*?)\s*(wx_\w+)\(([^)]*)\);", body, flags=re.M | re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        params = []
        if args != "void":
            for a in args.split(","):
                a = a.strip()
                am = re.match(r"(.*?)(\w+)(\[\d+\])?$", a)
                ctype, pname, arr = am.group(1).strip(), am.group(2), am.group(3)
                params.append((pname, c_type_to_rust(ctype, arr[1:-1] if arr else None, as_param=True)))
        out[name] = (params, c_type_to_rust(ret))
    return out


def rust_functions():
    out = {}
    block = re.search(r'extern "C" \{(.*?)\n\}', RS, flags=re.S).group(1)
    for m in re.finditer(r"pub fn (wx_\w+)\((.*?)\)\s*->\s*([^;]+);", block, flags=re.S):
        params = []
        for a in m.group(2).split(","):
            a = " ".join(a.split())
            if not a:
                continue
            pname, ptype = a.split(":", 1)
            params.append((pname.strip(), ptype.strip()))
        out[m.group(1)] = (params, m.group(3).strip())
    return out


def test_repr_c_structs_match_the_header():
    c, r = c_structs(), rust_structs()
    assert set(c) >= {"WxTreeDesc", "WxState", "WxAov", "WxShard", "WxTreeInfo", "WxRenderInfo", "WxSdfInfo"}
    for name, fields in c.items():
        assert name in r, f"{name} missing from the Rust crate"
        assert r[name] == fields, f"{name}: rust {r[name]} != header {fields}"


def test_wxstate_is_the_references_compute_state():
    """Field order and sizes of ComputeState (compute_state.rs:9-29): 2 x mat4, 4 x vec4 f32, 2 x vec4 u32, 2 x vec4 f32 = 256 B."""
    f = dict(rust_structs()["WxState"])
    assert list(f) == ["view_proj", "camera_to_world", "eye", "u", "mv", "wp", "render_mode", "show_345", "sun_dir", "sun_color"]
    size = sum({"f32": 4, "u32": 4}[re.match(r"\[(\w+); (\d+)\]", t).group(1)] * int(re.match(r"\[(\w+); (\d+)\]", t).group(2)) for t in f.values())
    assert size == 256


def test_extern_block_declares_every_function_of_the_header():
    c, r = c_functions(), rust_functions()
    from woxel_b200 import _ffi
    assert set(c) == set(_ffi.CUDA_API), set(c) ^ set(_ffi.CUDA_API)  # the ctypes table the symbol-export test checks
    assert set(c) == set(r), (sorted(set(c) - set(r)), sorted(set(r) - set(c)))
    for name, (params, ret) in c.items():
        rp, rr = r[name]
        assert rr == ret, f"{name}: return {rr} != {ret}"
        assert [t for _, t in rp] == [t for _, t in params], f"{name}: rust {rp} != header {params}"
        assert [n for n, _ in rp] == [n for n, _ in params], f"{name}: parameter names differ: {rp} vs {params}"


def test_constants_match_the_header():
    for m in re.finditer(r"(WX_(?:OK|ERR_\w+|OPT_\w+)) = (-?\d+)", strip_comments(HDR)):
        rm = re.search(rf"pub const {m.group(1)}: c_int = (-?\d+);", RS)
        assert rm and rm.group(1) == m.group(2), m.group(1)
    assert re.search(r"#define WX_ABI_VERSION (\d+)", HDR).group(1) == re.search(r"pub const WX_ABI_VERSION: c_int = (\d+);", RS).group(1)


def test_safe_crate_uses_only_declared_functions():
    safe = open(os.path.join(ROOT, "rust", "woxel-b200", "src", "lib.rs")).read()
    used = set(re.findall(r"\b(wx_\w+)\(", safe))
    assert used and used <= set(rust_functions()), used - set(rust_functions())
