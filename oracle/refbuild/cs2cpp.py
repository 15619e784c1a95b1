#!/usr/bin/env python3
"""cs2cpp.py — TEST INFRASTRUCTURE ONLY (recipe of oracle/_ref).

Rewrites the reference's own C# sources (pipliz/cpuvox, read where they lie under /root/reference and never copied into
this repository) into C++ so that g++ can compile them: the image has no C# toolchain (dotnet / mono / mcs / csc all absent,
also on the GPU box), so "the reference compiled here" is only reachable this way.  The rewrite is syntactic — tokens are
mapped, no expression is re-derived:

  * types are flattened (`Outer.Inner` -> `Outer_Inner`), members declared in-class and defined out of class, so C#'s
    order-free declarations survive C++'s single pass;
  * `ref`/`out` parameters become references (inline `out T x` declarations are hoisted in front of their statement);
  * properties become nullary methods and their uses gain `()`, swizzles (`.xz`) likewise;
  * `new T { a = x }` becomes `cs_init(T(), [&](T& _o){ _o.a = x; })`, `new T[n]` a shared array, `stackalloc` a zeroed alloca;
  * local functions become lambdas placed in front of their first use; `try/finally` becomes two blocks;
  * `1f` -> `1.f`, `float.Epsilon` -> the denormal constant, `null` -> `nullptr`, `default` -> `{}`, `X.Y` -> `X::Y` for types.

The Unity API the sources call (Unity.Mathematics, Collections, Jobs, UnityEngine) comes from unity_shim.hpp — our
restatement of those closed/un-vendored packages.  Output goes to the directory given with -o (oracle/_ref/, git-ignored).

usage: cs2cpp.py --ref /root/reference -o oracle/_ref/gen/ref_gen.hpp
"""
from __future__ import annotations

import argparse
import os
import re
import sys

# ------------------------------------------------------------------------------------------------------------------
# What is translated: file -> members left out (UI / Unity object plumbing that the path does not touch).
# A name excludes every overload; "Type.member" addresses members of nested types.
# ------------------------------------------------------------------------------------------------------------------
FILES = [
    ("Assets/Code/Utils/Color24.cs", {"ColorARGB32": ["operator ==", "operator !=", "Equals", "GetHashCode"]}),
    ("Assets/Code/UnityManager.cs", {"UnityManager": "ONLY:LOD_LEVELS,ERenderMode"}),
    ("Assets/Code/Utils/SegmentDDAData.cs", {}),
    ("Assets/Code/Utils/SimpleMesh.cs", {"SimpleMesh": "ONLY:Vertices,Indices,VertexCount,IndexCount,MallocHelper,FreeHelper,Dispose,Remap_Internal,Vertex"}),
    ("Assets/Code/VoxelizerHelper.cs", {"VoxelizerHelper": ["ExecuteDelegate", "GetVoxelsInvoker", "Initialize", "GetVoxels"]}),
    ("Assets/Code/World.cs", {}),
    ("Assets/Code/WordBuilder.cs", {"WorldBuilder": ["Import"]}),
    ("Assets/Code/Utils/CameraData.cs", {}),
    ("Assets/Code/Rendering/RayBuffer.cs", {}),
    ("Assets/Code/RenderManager.cs", {"RenderManager": ["ClearRayBuffer"]}),
    ("Assets/Code/Rendering/DrawSegmentRayJob.cs", {}),
    ("Assets/Code/WorldSaveFile.cs", {}),
]

# reference (class) types: variables of these types are references in the C++ text
CLASS_TYPES = {"RayBuffer", "Camera", "Transform", "SimpleMesh", "WorldBuilder", "Texture2D", "RenderTexture", "Mesh", "Material",
               "CommandBuffer", "RenderManager"}
# names the shim defines as types / static classes (for `X.Y` -> `X::Y`)
SHIM_TYPES = {
    "Allocator", "NativeArrayOptions", "UnsafeUtility", "Mathf", "Vector2", "Vector3", "Vector4", "Matrix4x4", "Quaternion",
    "Debug", "Profiler", "Color", "Color32", "float4x4", "math", "TextureFormat", "RenderTextureFormat", "FilterMode", "Object",
    "Interlocked", "Environment", "Parallel", "Texture2D", "Screen", "MeshTopology", "CameraEvent", "MemoryMappedFile", "FileMode",
}
SHIM_PROPERTIES = {"Count"}
KEYWORDS = {"if", "else", "while", "for", "foreach", "switch", "return", "new", "do", "try", "catch", "finally", "lock", "using",
            "fixed", "throw", "goto", "case", "break", "continue", "await", "yield", "in", "is", "as"}
MODIFIERS = ["public", "private", "internal", "protected", "unsafe", "readonly", "sealed", "partial", "override", "virtual", "extern", "new"]


def die(msg):
    sys.stderr.write("cs2cpp: " + msg + "\n")
    sys.exit(1)


# ------------------------------------------------------------------------------------------------------------------
# lexical helpers
# ------------------------------------------------------------------------------------------------------------------
def strip_comments(src: str) -> str:
    out, i, n = [], 0, len(src)
    while i < n:
        c = src[i]
        if c == '"' or (c == "$" and src[i + 1:i + 2] == '"') or (c == "@" and src[i + 1:i + 2] == '"'):
            j = i + (2 if c in "$@" else 1)
            while j < n and src[j] != '"':
                j += 2 if src[j] == "\\" else 1
            out.append(src[i:j + 1])
            i = j + 1
        elif c == "'":
            j = i + 1
            while j < n and src[j] != "'":
                j += 2 if src[j] == "\\" else 1
            out.append(src[i:j + 1])
            i = j + 1
        elif src.startswith("//", i):
            j = src.find("\n", i)
            i = n if j < 0 else j
        elif src.startswith("/*", i):
            j = src.find("*/", i)
            i = n if j < 0 else j + 2
        else:
            out.append(c)
            i += 1
    return "".join(out)


OPEN = {"(": ")", "[": "]", "{": "}"}


def match(text: str, i: int) -> int:
    """index of the bracket closing text[i]"""
    stack = [OPEN[text[i]]]
    j = i + 1
    n = len(text)
    while j < n:
        c = text[j]
        if c == '"':
            j += 1
            while text[j] != '"':
                j += 2 if text[j] == "\\" else 1
        elif c == "'":
            j += 1
            while text[j] != "'":
                j += 2 if text[j] == "\\" else 1
        elif c in OPEN:
            stack.append(OPEN[c])
        elif c in ")]}":
            if c != stack.pop():
                die("bracket mismatch near: " + text[max(0, j - 60):j + 20])
            if not stack:
                return j
        j += 1
    die("unbalanced bracket from: " + text[i:i + 80])


def match_angle(text: str, i: int) -> int:
    depth, j = 0, i
    while j < len(text):
        if text[j] == "<":
            depth += 1
        elif text[j] == ">":
            depth -= 1
            if depth == 0:
                return j
        elif text[j] in ";{}()":
            return -1
        j += 1
    return -1


def split_top(text: str, sep: str = ",") -> list[str]:
    parts, depth, cur, i = [], 0, [], 0
    while i < len(text):
        c = text[i]
        if c in "([{":
            j = match(text, i)
            cur.append(text[i:j + 1])
            i = j + 1
            continue
        if c == "<":
            j = match_angle(text, i)
            if j > 0 and re.match(r"[\w\s,.<>\[\]*?]*$", text[i + 1:j]):
                cur.append(text[i:j + 1])
                i = j + 1
                continue
        if c == sep and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(c)
        i += 1
    parts.append("".join(cur))
    return parts


ATTR_RE = re.compile(r"(?<=[\s])\[\s*(?:[A-Za-z_][\w.]*)\s*(?:\((?:[^()\[\]]|\([^()]*\))*\))?\s*(?:,\s*[A-Za-z_][\w.]*\s*(?:\([^()]*\))?\s*)*\](?=\s)")


def preprocess(src: str) -> str:
    src = src.lstrip("﻿")
    src = strip_comments(src)
    src = re.sub(r"^\s*using\s+[\w.= ]+;\s*$", "", src, flags=re.M)
    src = re.sub(r"^\s*#(pragma|region|endregion).*$", "", src, flags=re.M)
    prev = None
    while prev != src:
        prev = src
        src = ATTR_RE.sub("", src)
    return src


# ------------------------------------------------------------------------------------------------------------------
# declarations
# ------------------------------------------------------------------------------------------------------------------
class Member:
    def __init__(self, kind, name, header, body=None, expr=None):
        self.kind, self.name, self.header, self.body, self.expr = kind, name, header, body, expr


class TypeDecl:
    def __init__(self, kind, name, generics, bases, outer):
        self.kind, self.name, self.generics, self.bases, self.outer = kind, name, generics, bases, outer
        self.members: list[Member] = []
        self.nested: list[TypeDecl] = []
        self.is_static = False
        self.enum_body = None

    @property
    def qual(self):
        return (self.outer.qual + "." if self.outer else "") + self.name

    @property
    def mangled(self):
        return self.qual.replace(".", "_")


def strip_mods(header: str):
    mods = set()
    changed = True
    while changed:
        changed = False
        header = header.lstrip()
        for m in MODIFIERS + ["static", "const", "fixed"]:
            if re.match(m + r"\b", header):
                mods.add(m)
                header = header[len(m):]
                changed = True
    return mods, header.strip()


def parse_members(text: str, owner: TypeDecl):
    i, n = 0, len(text)
    while True:
        while i < n and text[i].isspace():
            i += 1
        if i >= n:
            return
        start, j, brace = i, i, -1
        while j < n:
            c = text[j]
            if c in "([":
                j = match(text, j)
            elif c == "{":
                brace = j
                break
            elif c == ";":
                break
            elif c == '"':
                j += 1
                while text[j] != '"':
                    j += 2 if text[j] == "\\" else 1
            j += 1
        if j >= n:
            die("member without end in " + owner.qual + ": " + text[start:start + 80])
        head = text[start:j]
        if brace >= 0 and re.search(r"(?<![=!<>])=(?![=>])", re.sub(r"\([^()]*\)", "", head)) and "(" not in head.split("=")[0]:
            # field with a brace initialiser: runs to the ';'
            k = match(text, brace)
            k = text.index(";", k)
            add_member(owner, text[start:k], None)
            i = k + 1
            continue
        if brace < 0:
            add_member(owner, head, None)
            i = j + 1
            continue
        k = match(text, brace)
        add_member(owner, head, text[brace + 1:k])
        i = k + 1


def add_member(owner: TypeDecl, header: str, body):
    mods, h = strip_mods(header)
    h = re.sub(r"\s+", " ", h).strip()
    if not h:
        return
    m = re.match(r"(struct|class|interface|enum)\s+(\w+)\s*(<[^>]*>)?\s*(?::\s*([^{]+?))?\s*(?:where .*)?$", h)
    if m and body is not None:
        t = TypeDecl(m.group(1), m.group(2), m.group(3), [b.strip() for b in (m.group(4) or "").split(",") if b.strip()], owner)
        t.is_static = "static" in mods
        if t.kind == "enum":
            t.enum_body = body
        else:
            parse_members(body, t)
        owner.nested.append(t)
        return
    if h.startswith("delegate "):
        owner.members.append(Member("delegate", h.split("(")[0].split()[-1], h))
        return
    mem = None
    if "=>" in h and body is None:
        left, expr = h.split("=>", 1)
        left = left.strip()
        if "(" in left:
            mem = Member("method", re.match(r".*?(\w+)\s*(<[^>]*>)?\s*\(", left).group(1), left, body="return " + expr.strip() + ";")
        else:
            mem = Member("property", left.split()[-1], left, body="get { return " + expr.strip() + "; }")
    elif body is not None and "(" not in h.split("=")[0] and " this[" not in h and not re.search(r"\boperator\b", h):
        mem = Member("property", h.split()[-1], h, body=body)
    elif body is not None and " this[" in h:
        mem = Member("indexer", "this[]", h, body=body)
    elif "(" in h and (body is not None) :
        om = re.match(r"(?:implicit|explicit) operator\s+([\w.]+)\s*\(", h)
        if om:
            mem = Member("convop", "operator " + om.group(1), h, body=body)
        else:
            om = re.match(r".*?\boperator\s*(\S+?)\s*\(", h)
            if om:
                mem = Member("operator", "operator " + om.group(1), h, body=body)
            else:
                nm = re.match(r"(.*?)(\w+)\s*(<[^>]*>)?\s*\(", h)
                name = nm.group(2)
                kind = "ctor" if (name == owner.name and nm.group(1).strip() == "") else "method"
                mem = Member(kind, name, h, body=body)
    else:
        # field (maybe several declarators, maybe an initialiser)
        decl = h.split("=")[0].strip()
        name = re.sub(r"\[.*?\]", "", decl).split(",")[0].split()[-1]
        mem = Member("field", name, h)
    mem.mods = mods
    owner.members.append(mem)


def parse_file(src: str) -> list[TypeDecl]:
    root = TypeDecl("root", "", None, [], None)
    parse_members(preprocess(src), root)
    for t in root.nested:
        t.outer = None
    return root.nested


# ------------------------------------------------------------------------------------------------------------------
# translation context
# ------------------------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self):
        self.types: dict[str, TypeDecl] = {}  # qualified C# name -> decl
        self.properties: set[str] = set(SHIM_PROPERTIES)
        self.auto_properties: set[tuple[str, str]] = set()

    def all_types(self):
        return list(self.types.values())

    def type_names(self):
        return {t.name for t in self.types.values()} | SHIM_TYPES


def walk(t: TypeDecl):
    yield t
    for n in t.nested:
        yield from walk(n)


def apply_selection(t: TypeDecl, sel):
    spec = sel.get(t.name)
    if spec is None:
        return
    if isinstance(spec, str) and spec.startswith("ONLY:"):
        keep = set(spec[5:].split(","))
        t.members = [m for m in t.members if m.name in keep]
        t.nested = [n for n in t.nested if n.name in keep]
    else:
        drop = set(spec)
        t.members = [m for m in t.members if m.name not in drop]
        t.nested = [n for n in t.nested if n.name not in drop]
    for n in t.nested:
        apply_selection(n, sel)


# ------------------------------------------------------------------------------------------------------------------
# body rewriting
# ------------------------------------------------------------------------------------------------------------------
def conv_type(ty: str, ctx: Ctx, scope: TypeDecl | None) -> str:
    """C# type text -> C++ type text (mangled nested names, arrays, class references handled by the caller)"""
    ty = ty.strip()
    ty = re.sub(r"\b(?:Unity\.Mathematics|System\.Threading\.Tasks|System\.Threading|System\.Collections\.Generic|System\.IO|Unity\.Collections|UnityEngine\.Rendering|UnityEngine)\.", "", ty)
    ty = qualify_types(ty, ctx, scope)
    ty = re.sub(r"(\b\w+<[\w:,\s*]+>)\.(?=[A-Z])", r"\1::", ty)
    # T[] -> ManagedArray<T>
    prev = None
    while prev != ty:
        prev = ty
        ty = re.sub(r"([\w:]+(?:<[^<>\[\]]*>)?\**)\s*\[\]", r"ManagedArray<\1>", ty)
    return ty


def qualify_types(text: str, ctx: Ctx, scope: TypeDecl | None) -> str:
    """`Outer.Inner` -> Outer_Inner everywhere; bare `Inner` -> mangled inside the scope chain that declares it"""
    quals = sorted((q for q in ctx.types if "." in q), key=len, reverse=True)
    for q in quals:
        text = re.sub(r"(?<![\w.])" + re.escape(q) + r"\b", ctx.types[q].mangled, text)
    s = scope
    seen = set()
    while s is not None:
        # nested types of s are visible unqualified inside s (and inside its nested types)
        for n in s.nested:
            if n.name not in seen:
                seen.add(n.name)
                text = re.sub(r"(?<![\w.:>])" + n.name + r"\b(?!\s*::)", n.mangled, text)
        s = s.outer
    # partially qualified names (`Inner.Innermost` seen from inside Outer): join onto already mangled prefixes
    mangled = {t.mangled for t in ctx.types.values()}
    prev = None
    while prev != text:
        prev = text
        text = re.sub(r"\b(\w+)\.(\w+)\b", lambda m: m.group(1) + "_" + m.group(2) if (m.group(1) in mangled and m.group(1) + "_" + m.group(2) in mangled) else m.group(0), text)
    return text


def hoist_local_functions(body: str, ctx: Ctx, scope) -> str:
    """local functions -> lambdas in front of their first use (recursively for nested blocks)"""
    # find local function definitions at depth 0 of this block
    i, n = 0, len(body)
    funcs = []  # (start, end, ret, name, params, inner)
    stmt_starts = [0]
    pat = re.compile(r"\s*(?:static\s+)?([\w.<>\[\]*]+)\s+(\w+)\s*\(([^()]*)\)\s*\{")
    out = []
    last = 0
    while i < n:
        c = body[i]
        at_start = i == 0 or body[:i].rstrip()[-1:] in (";", "{", "}", "")
        if at_start:
            m = pat.match(body, i)
            if m and m.group(1) not in KEYWORDS and m.group(2) not in KEYWORDS and m.group(1) != "new":
                b = m.end() - 1
                e = match(body, b)
                funcs.append((m.group(1), m.group(2), m.group(3), body[b + 1:e]))
                out.append(body[last:i])
                last = e + 1
                i = e + 1
                continue
        if c in "([":
            i = match(body, i) + 1
            continue
        if c == "{":
            e = match(body, i)
            inner = hoist_local_functions(body[i + 1:e], ctx, scope)
            out.append(body[last:i + 1] + inner + "}")
            last = e + 1
            i = e + 1
            continue
        if c == '"':
            i += 1
            while body[i] != '"':
                i += 2 if body[i] == "\\" else 1
        i += 1
    out.append(body[last:])
    text = "".join(out)
    if not funcs:
        return text
    # order: callee before caller
    names = [f[1] for f in funcs]
    ordered, pending = [], list(funcs)
    while pending:
        for f in pending:
            deps = [g for g in pending if g is not f and re.search(r"\b" + g[1] + r"\s*\(", f[3])]
            if not deps:
                ordered.append(f)
                pending.remove(f)
                break
        else:
            die("recursive local functions")
    lambdas = ""
    for ret, name, params, inner in ordered:
        inner = hoist_local_functions(inner, ctx, scope)
        lambdas += "auto %s = [&](%s) -> %s {%s};\n" % (name, conv_params(params, ctx, scope, defaults=False), conv_type(ret, ctx, scope), inner)
    # first use
    first = min((m.start() for nm in names for m in [re.search(r"\b" + nm + r"\s*\(", text)] if m), default=None)
    if first is None:
        return text
    # statement start at depth 0 before `first`
    pos, i, depth0_starts = 0, 0, [0]
    while i < first:
        c = text[i]
        if c in "([{":
            e = match(text, i)
            if e >= first:
                break
            i = e + 1
            if c == "{":
                rest = text[i:].lstrip()
                if not re.match(r"(else|catch|finally|while\b[^;{]*;)", rest) and not rest.startswith((";", ",", ")")):
                    depth0_starts.append(i)
            continue
        if c == ";":
            depth0_starts.append(i + 1)
        i += 1
    pos = depth0_starts[-1]
    return text[:pos] + "\n" + lambdas + text[pos:]


def conv_params(params: str, ctx: Ctx, scope, defaults=True) -> str:
    out = []
    for p in split_top(params):
        p = p.strip()
        if not p:
            continue
        default = None
        if "=" in p:
            p, default = [s.strip() for s in p.split("=", 1)]
        byref = False
        m = re.match(r"(ref|out|in|this)\s+(.*)", p)
        if m:
            byref = m.group(1) in ("ref", "out")
            p = m.group(2)
        ty, name = p.rsplit(None, 1)
        cty = conv_type(ty, ctx, scope)
        if byref or ty.strip().split(".")[-1] in CLASS_TYPES:
            cty += "&"
        s = cty + " " + name
        if default is not None and defaults:
            s += " = " + conv_expr_text(default, ctx, scope)
        out.append(s)
    return ", ".join(out)


def conv_new(text: str, ctx: Ctx, scope) -> str:
    """new T(args) {inits} / new T[n] / new T[] {..}"""
    out, i = [], 0
    pat = re.compile(r"\bnew\s+([A-Za-z_][\w.]*)")
    while True:
        m = pat.search(text, i)
        if not m:
            out.append(text[i:])
            break
        out.append(text[i:m.start()])
        j = m.end()
        ty = m.group(1)
        # generics
        k = j
        while k < len(text) and text[k].isspace():
            k += 1
        if k < len(text) and text[k] == "<":
            e = match_angle(text, k)
            ty += text[k:e + 1]
            j = e + 1
        while j < len(text) and text[j] in "?*":
            ty += text[j]
            j += 1
        k = j
        while k < len(text) and text[k].isspace():
            k += 1
        cty = conv_type(ty, ctx, scope)
        if k < len(text) and text[k] == "[":
            e = match(text, k)
            size = text[k + 1:e].strip()
            k2 = e + 1
            while k2 < len(text) and text[k2].isspace():
                k2 += 1
            if size == "" and k2 < len(text) and text[k2] == "{":
                e2 = match(text, k2)
                out.append("ManagedArray<%s>{%s}" % (cty, conv_new(text[k2 + 1:e2], ctx, scope)))
                i = e2 + 1
            else:
                out.append("ManagedArray<%s>(%s)" % (cty, conv_new(size, ctx, scope)))
                i = e + 1
            continue
        args = ""
        if k < len(text) and text[k] == "(":
            e = match(text, k)
            args = conv_new(text[k + 1:e], ctx, scope)
            j = e + 1
            k = j
            while k < len(text) and text[k].isspace():
                k += 1
        ctor = "%s(%s)" % (cty, args)
        if ty.split("<")[0] in CLASS_TYPES:
            ctor = "cs_new(%s)" % ctor
        if ty.startswith("List<"):
            ctor = "%s::Create()" % cty
        if k < len(text) and text[k] == "{":
            e = match(text, k)
            inits = ""
            for part in split_top(text[k + 1:e]):
                part = part.strip()
                if not part:
                    continue
                nm, val = part.split("=", 1)
                inits += " _o.%s = %s;" % (nm.strip(), conv_new(val.strip(), ctx, scope))
            out.append("cs_init(%s, [&](%s& _o) {%s })" % (ctor, cty, inits))
            i = e + 1
        else:
            out.append(ctor)
            i = j
    return "".join(out)


def hoist_out_vars(text: str, ctx: Ctx, scope) -> str:
    pat = re.compile(r"\bout\s+((?:[A-Za-z_][\w.]*)(?:<[^<>()]*>)?\**)\s+(\w+)\s*(?=[,)])")
    while True:
        m = pat.search(text)
        if not m:
            return text
        ty, name = m.group(1), m.group(2)
        # start of the enclosing statement
        j = m.start()
        depth = 0
        while j > 0:
            c = text[j - 1]
            if c in ")]":
                depth += 1
            elif c in "([":
                depth -= 1
            elif c in ";{}" and depth <= 0:
                break
            j -= 1
        decl = "%s %s;\n" % (conv_type(ty, ctx, scope), name)
        text = text[:j] + "\n" + decl + text[j:m.start()] + name + text[m.end():]


def conv_try(text: str) -> str:
    out, i = [], 0
    pat = re.compile(r"\btry\s*\{")
    while True:
        m = pat.search(text, i)
        if not m:
            out.append(text[i:])
            return "".join(out)
        b = m.end() - 1
        e = match(text, b)
        rest = text[e + 1:]
        inner = conv_try(text[b + 1:e])
        fm = re.match(r"\s*finally\s*\{", rest)
        if fm:
            fb = e + 1 + fm.end() - 1
            fe = match(text, fb)
            out.append(text[i:m.start()] + "{" + inner + "} /*finally*/ {" + conv_try(text[fb + 1:fe]) + "}")
            i = fe + 1
        else:
            out.append(text[i:m.start()] + "try {" + inner + "}")
            i = e + 1


def conv_expr_text(text: str, ctx: Ctx, scope, owner_statics=None) -> str:
    """token level rewrites shared by bodies, initialisers and default values"""
    text = re.sub(r"\b(?:Unity\.Mathematics|System\.Threading\.Tasks|System\.Threading|System\.Collections\.Generic|System\.IO|Unity\.Collections|UnityEngine\.Rendering|UnityEngine\.Profiling|UnityEngine)\.", "", text)
    text = conv_new(text, ctx, scope)
    text = hoist_out_vars(text, ctx, scope)
    text = conv_try(text)
    # ref locals, ref/out arguments
    text = re.sub(r"\bref\s+([\w.<>]+)\s+(\w+)\s*=\s*ref\s+", lambda m: conv_type(m.group(1), ctx, scope) + "& " + m.group(2) + " = ", text)
    text = re.sub(r"\bref\s+this\b", "*this", text)
    text = re.sub(r"(?<=[(,])\s*(?:ref|out)\s+", " ", text)
    text = re.sub(r"=\s*this\s*;", "= *this;", text)
    text = re.sub(r"\bthis\s*=\s*default\s*;", "*this = {};", text)
    text = re.sub(r"\bthis\.", "this->", text)
    text = re.sub(r"\(\s*this\s*\)", "(*this)", text)
    text = re.sub(r"\bnull\b", "nullptr", text)
    text = re.sub(r"(=|return|\?|:)\s*default\s*(?=[;,)])", r"\1 {}", text)
    text = re.sub(r"\bdefault\s*\(\s*([\w.]+)\s*\)", lambda m: conv_type(m.group(1), ctx, scope) + "{}", text)
    # literals
    text = re.sub(r"(?<![\w.])(\d+)[fF]\b", r"\1.f", text)
    text = re.sub(r"\bfloat\.(Epsilon|NegativeInfinity|PositiveInfinity)\b", r"cs_float_\1", text)
    text = re.sub(r"\bint\.(MaxValue|MinValue)\b", r"cs_int_\1", text)
    text = re.sub(r"\bstackalloc\s+(\w+)\s*\[([^\]]+)\]", r"cs_stackalloc(\1, \2)", text)
    # statements
    text = re.sub(r"\bfixed\s*\(\s*([\w*.]+\s+\w+)\s*=\s*([\w.]+)\s*\)\s*\{", r"{ \1 = (\2).data();", text)
    text = re.sub(r"\busing\s*\(\s*((?:[^()]|\((?:[^()]|\([^()]*\))*\))*)\)\s*\{", r"{ \1;", text)   # using (T x = ...) { -> scoped object
    text = re.sub(r"\bforeach\s*\(\s*(?:var|[\w.<>]+)\s+(\w+)\s+in\s+", r"for (auto& \1 : ", text)
    text = re.sub(r"\bvar\b", "auto", text)
    text = re.sub(r"\block\s*\([^()]*\)\s*\{", "{", text)
    text = re.sub(r"\bthrow\s+new\s+", "throw ", text)
    text = re.sub(r"\bcatch\s*\(\s*(?:System\.)?Exception\s+(\w+)\s*\)", r"catch (std::exception& \1)", text)
    # lambdas: (int i) => { ... }  /  name => { ... }  / (a, b) => expr
    text = re.sub(r"\(\s*(int\s+\w+)\s*\)\s*=>\s*(?=\{)", r"[&](\1) ", text)
    text = re.sub(r"(?<=[(,])\s*(\w+)\s*=>\s*(?=\{)", r" [&](auto \1) ", text)
    text = re.sub(r"\(\s*\)\s*=>\s*(?=\{)", r"[&]() ", text)
    text = conv_expr_lambdas(text)
    text = re.sub(r"(\w+(?:\.\w+)*)\.CompareTo\(([^()]*)\)", r"cs_compare(\1, \2)", text)
    # swizzles
    text = re.sub(r"\.([xyzw]{2,4})\b(?!\s*\()", r".\1()", text)
    # types
    text = qualify_types(text, ctx, scope)
    prev = None
    while prev != text:
        prev = text
        text = re.sub(r"\b([A-Za-z_][\w:]*(?:<[^<>\[\]();]*>)?\**)\s*\[\](?=\s+\w)", r"ManagedArray<\1>", text)
    # class-typed locals become references
    for c in CLASS_TYPES:
        text = re.sub(r"(?<![\w.:<&*])" + c + r"\s+(\w+)\s*(?==[^=])", c + r"& \1 ", text)
        text = re.sub(c + r"& (\w+) =\s*cs_new\(", c + r" \1 = cs_new(", text)  # a fresh object is held by value
    # X.Y -> X::Y for types and static classes, generic nested types
    names = ctx.type_names() | {t.mangled for t in ctx.types.values()}
    text = re.sub(r"(?<![\w.>])(" + "|".join(sorted(map(re.escape, names), key=len, reverse=True)) + r")\.(?=[A-Za-z_])", r"\1::", text)
    text = re.sub(r"(\b\w+<[\w:,\s*]+>)\.(?=[A-Z])", r"\1::", text)
    # properties -> calls
    if ctx.properties:
        text = re.sub(r"(?<![\w:])(" + "|".join(sorted(ctx.properties, key=len, reverse=True)) + r")\b(?!\s*[(\w])", r"\1()", text)
    # statics of the enclosing type, seen from a flattened nested type
    if owner_statics:
        for outer_mangled, names_ in owner_statics:
            if names_:
                text = re.sub(r"(?<![\w.:>])(" + "|".join(sorted(names_, key=len, reverse=True)) + r")\b(?!\s*::)", outer_mangled + r"::\1", text)
    return text


def conv_expr_lambdas(text: str) -> str:
    """(a, b) => expr   ->   [&](auto a, auto b) { return expr; }   (expression ends at the call's closing bracket)"""
    pat = re.compile(r"\(\s*(\w+)\s*,\s*(\w+)\s*\)\s*=>\s*(?!\{)")
    while True:
        m = pat.search(text)
        if not m:
            return text
        j, depth = m.end(), 0
        while j < len(text):
            c = text[j]
            if c in "([{":
                j = match(text, j)
            elif c in ")]};" or (c == "," and depth == 0):
                break
            j += 1
        text = text[:m.start()] + "[&](auto %s, auto %s) { return %s; }" % (m.group(1), m.group(2), text[m.end():j].strip()) + text[j:]


def conv_body(body: str, ctx: Ctx, scope, owner_statics) -> str:
    body = hoist_local_functions(body, ctx, scope)
    return conv_expr_text(body, ctx, scope, owner_statics)


# ------------------------------------------------------------------------------------------------------------------
# emission
# ------------------------------------------------------------------------------------------------------------------
def statics_visible_from(t: TypeDecl):
    """(mangled outer, names) for every enclosing type: its static methods / constants / static fields, minus names t declares"""
    own = {m.name for m in t.members} | {n.name for n in t.nested}
    res = []
    o = t.outer
    while o is not None:
        names = set()
        for m in o.members:
            if m.kind in ("method", "field") and (o.is_static or "static" in m.mods or "const" in m.mods) and m.name not in own:
                names.add(m.name)
        res.append((o.mangled, names))
        own |= names
        o = o.outer
    return res


def field_decl(m: Member, ctx: Ctx, scope: TypeDecl, in_static_class: bool) -> str:
    h = m.header
    init = None
    if "=" in h:
        h, init = [s.strip() for s in h.split("=", 1)]
    fixed = re.match(r"(.*?)\s+(\w+)\s*\[(.*)\]$", h)
    if "fixed" in m.mods and fixed:
        ty = conv_type(fixed.group(1), ctx, scope)
        return "%s %s[%s] = {};" % (ty, fixed.group(2), conv_expr_text(fixed.group(3), ctx, scope))
    parts = split_top(h)
    ty, first = parts[0].rsplit(None, 1)
    cty = conv_type(ty, ctx, scope)
    prefix = ""
    if "const" in m.mods:
        prefix = "static constexpr "
    elif "static" in m.mods or in_static_class:
        prefix = "static inline "
    names = [first] + [p.strip() for p in parts[1:]]
    if init is not None:
        return "%s%s %s = %s;" % (prefix, cty, ", ".join(names), conv_expr_text(init, ctx, scope))
    if len(names) == 1 and not prefix:
        if names[0] in ctx.properties:
            # a field whose name is a property elsewhere: uses were rewritten to `Name()`, so give it an accessor
            return "%s %s_field = {}; %s& %s() { return %s_field; }" % (cty, names[0], cty, names[0], names[0])
        return "%s %s = {};" % (cty, names[0])
    return "%s%s %s;" % (prefix, cty, ", ".join(names))


def emit_type_decl(t: TypeDecl, ctx: Ctx) -> str:
    if t.kind == "enum":
        return "enum class %s { %s };\n" % (t.mangled, t.enum_body.strip())
    bases = []
    for b in t.bases:
        if b.startswith("IJobParallelFor"):
            bases.append("IJobParallelFor<%s>" % t.mangled)
    lines = ["struct %s%s {" % (t.mangled, (" : " + ", ".join(bases)) if bases else "")]
    has_ctor = any(m.kind == "ctor" for m in t.members)
    has_default = any(m.kind == "ctor" and re.match(r"\w+\s*\(\s*\)", m.header) for m in t.members)
    if has_ctor and not has_default:
        lines.append("    %s() = default;" % t.mangled)
    for m in t.members:
        static = "static " if (t.is_static or "static" in m.mods) else ""
        if m.kind == "field":
            lines.append("    " + field_decl(m, ctx, t, t.is_static))
        elif m.kind == "ctor":
            params = m.header[m.header.index("(") + 1:match(m.header, m.header.index("("))]
            lines.append("    %s(%s);" % (t.mangled, conv_params(params, ctx, t)))
        elif m.kind == "method":
            ret, name, gen, params = split_method_header(m.header)
            tmpl = ("template <%s> " % ", ".join("class " + g.strip() for g in gen.split(","))) if gen else ""
            lines.append("    %s%s%s %s(%s);" % (tmpl, static, conv_type(ret, ctx, t), name, conv_params(params, ctx, t)))
        elif m.kind == "property":
            ty = m.header.rsplit(None, 1)[0]
            cty = conv_type(ty, ctx, t)
            if re.fullmatch(r"\s*get\s*;\s*(?:(?:private\s+)?set\s*;)?\s*", m.body):
                lines.append("    %s %s_auto = {};" % (cty, m.name))
                lines.append("    %s%s& %s() { return %s_auto; }" % (static, cty, m.name, m.name))
            else:
                lines.append("    %s%s %s();" % (static, cty, m.name))
        elif m.kind == "convop":
            target = m.name.split()[1]
            src_param = m.header[m.header.index("(") + 1:match(m.header, m.header.index("("))]
            if target == t.name:
                lines.append("    %s(%s);" % (t.mangled, conv_params(src_param, ctx, t)))
            else:
                lines.append("    operator %s() const;" % conv_type(target, ctx, t))
    lines.append("};")
    return "\n".join(lines) + "\n"


def split_method_header(h: str):
    p = h.index("(")
    # generic method: Name<T> (
    m = re.match(r"(.*?)(\w+)\s*(?:<([^>]*)>)?\s*$", h[:p])
    ret, name, gen = m.group(1).strip(), m.group(2), m.group(3)
    params = h[p + 1:match(h, p)]
    return ret, name, gen, params


def emit_type_defs(t: TypeDecl, ctx: Ctx) -> str:
    if t.kind == "enum":
        return ""
    out = []
    statics = statics_visible_from(t)
    for m in t.members:
        if m.kind == "ctor":
            params = m.header[m.header.index("(") + 1:match(m.header, m.header.index("("))]
            out.append("inline %s::%s(%s) {%s}\n" % (t.mangled, t.mangled, conv_params(params, ctx, t, defaults=False), conv_body(m.body, ctx, t, statics)))
        elif m.kind == "method":
            ret, name, gen, params = split_method_header(m.header)
            tmpl = ("template <%s> " % ", ".join("class " + g.strip() for g in gen.split(","))) if gen else ""
            out.append("%sinline %s %s::%s(%s) {%s}\n" % (tmpl, conv_type(ret, ctx, t), t.mangled, name, conv_params(params, ctx, t, defaults=False), conv_body(m.body, ctx, t, statics)))
        elif m.kind == "property":
            if re.fullmatch(r"\s*get\s*;\s*(?:(?:private\s+)?set\s*;)?\s*", m.body):
                continue
            ty = m.header.rsplit(None, 1)[0]
            g = re.search(r"\bget\s*\{", m.body)
            if not g:
                die("property without getter: " + t.qual + "." + m.name)
            b = g.end() - 1
            e = match(m.body, b)
            out.append("inline %s %s::%s() {%s}\n" % (conv_type(ty, ctx, t), t.mangled, m.name, conv_body(m.body[b + 1:e], ctx, t, statics)))
        elif m.kind == "convop":
            target = m.name.split()[1]
            src_param = m.header[m.header.index("(") + 1:match(m.header, m.header.index("("))]
            body = conv_body(m.body, ctx, t, statics)
            if target == t.name:
                out.append("inline %s::%s(%s) { *this = [&]() -> %s {%s}(); }\n" % (t.mangled, t.mangled, conv_params(src_param, ctx, t, defaults=False), t.mangled, body))
            else:
                pname = src_param.split()[-1]
                out.append("inline %s::operator %s() const { const %s& %s = *this; %s}\n" % (t.mangled, conv_type(target, ctx, t), t.mangled, pname, body))
    return "".join(out)


HANDLE_TEMPLATES = re.compile(r"\b(?:NativeArray|NativeList|ManagedArray|List|NativeArrayList)\s*<[^<>]*(?:<[^<>]*>)?[^<>]*>(?:::\w+)?")


def by_value_deps(t: TypeDecl, ctx: Ctx) -> set[str]:
    deps = set()
    if t.kind == "enum":
        return deps
    mangled = {x.mangled for x in ctx.types.values()}
    for m in t.members:
        if m.kind != "field" and not (m.kind == "property" and re.fullmatch(r"\s*get\s*;\s*(?:set\s*;)?\s*", m.body or "x")):
            continue
        txt = field_decl(m, ctx, t, t.is_static) if m.kind == "field" else conv_type(m.header.rsplit(None, 1)[0], ctx, t)
        txt = HANDLE_TEMPLATES.sub("", txt)
        for tok in re.finditer(r"\b(\w+)\b(\s*\*)?", txt):
            if tok.group(1) in mangled and not tok.group(2) and tok.group(1) != t.mangled:
                deps.add(tok.group(1))
    return deps


def conv_shader(path: str) -> str:
    """default variant of RayBufferBlit.shader's fragment function (the text between `#else` and `#endif` inside `frag`)"""
    with open(path, encoding="utf-8") as f:
        src = strip_comments(f.read().lstrip("\ufeff"))
    m = re.search(r"fixed4\s+frag\s*\(\s*v2f\s+i\s*\)\s*:\s*SV_Target\s*\{(.*?)#else(.*?)#endif", src, flags=re.S)
    if not m:
        die("fragment function not found in " + path)
    body = m.group(2)
    body = re.sub(r"\b(_RayOffset|_RayScale)\s*\[([^\]]+)\]", r"g.\1[(int)(\2)]", body)
    body = re.sub(r"(?<![\w.])(_ScreenParams|_MainTex1|_MainTex2)\b", r"g.\1", body)
    return "inline float4 RayBufferBlit_frag(const v2f& i, const ShaderGlobals& g) {%s}\n" % body


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("-o", "--out", required=True)
    args = ap.parse_args()

    ctx = Ctx()
    tops: list[TypeDecl] = []
    for rel, sel in FILES:
        path = os.path.join(args.ref, rel)
        with open(path, encoding="utf-8") as f:
            decls = parse_file(f.read())
        for d in decls:
            apply_selection(d, sel)
            tops.append(d)
            for t in walk(d):
                ctx.types[t.qual] = t
    for t in ctx.all_types():
        for m in t.members:
            if m.kind == "property":
                ctx.properties.add(m.name)

    flat = ctx.all_types()
    by_mangled = {t.mangled: t for t in flat}
    # definition order: by-value field dependencies first
    order, done = [], set()

    def visit(t, stack=()):
        if t.mangled in done:
            return
        if t.mangled in stack:
            die("by-value cycle: " + " -> ".join(stack + (t.mangled,)))
        for d in sorted(by_value_deps(t, ctx)):
            visit(by_mangled[d], stack + (t.mangled,))
        done.add(t.mangled)
        order.append(t)

    for t in flat:
        visit(t)

    out = ["// GENERATED by oracle/refbuild/cs2cpp.py from the reference's C# sources — do not commit, do not edit.\n",
           "#pragma once\n#include \"unity_shim.hpp\"\n#include \"ref_prelude.hpp\"\nnamespace cpuvox_ref {\n"]
    for t in flat:
        if t.kind != "enum":
            out.append("struct %s;\n" % t.mangled)
    for t in order:
        if t.kind == "enum":
            out.append(emit_type_decl(t, ctx))
    for t in order:
        if t.kind != "enum":
            out.append(emit_type_decl(t, ctx))
    for t in order:
        out.append(emit_type_defs(t, ctx))
    out.append(conv_shader(os.path.join(args.ref, "Assets/Shaders/RayBufferBlit.shader")))
    out.append("}  // namespace cpuvox_ref\n")
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as f:
        f.write("".join(out))


if __name__ == "__main__":
    main()
