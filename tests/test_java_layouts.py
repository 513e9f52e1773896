"""java/.../VsrLayouts.java (Panama MemoryLayouts, source only: no JDK in this image) against the ctypes mirror of
include/vsr_b200.h: every struct, every field -- name, order, size, offset -- and the enum constants."""
import ctypes as C
import importlib
import os
import re

H = importlib.import_module("2dhmsr_b200")
A = H.abi

JAVA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "java", "it", "units", "erallab", "hmsrobots", "b200")
SIZES = {"JAVA_DOUBLE": 8, "JAVA_INT": 4, "JAVA_LONG": 8, "ADDRESS": 8, "JAVA_BYTE": 1}


def _parse_layouts(src):
    """{struct name: [(field, size)]} from the structLayout(...) blocks (one field per line)."""
    out, java_names = {}, {}
    for m in re.finditer(r"StructLayout (\w+) = MemoryLayout\.structLayout\((.*?)\)\.withName\(\"(\w+)\"\);", src, re.S):
        fields = []
        for line in m.group(2).strip().splitlines():
            line = line.strip().rstrip(",")
            f = re.fullmatch(r"(\w+)\.withName\(\"(\w+)\"\)", line)
            s = re.fullmatch(r"MemoryLayout\.sequenceLayout\((\d+), (\w+)\)\.withName\(\"(\w+)\"\)", line)
            p = re.fullmatch(r"MemoryLayout\.paddingLayout\((\d+)\)", line)
            if f and f.group(1) in SIZES:
                fields.append((f.group(2), SIZES[f.group(1)]))
            elif f:                                       # a nested struct declared above
                fields.append((f.group(2), sum(sz for _, sz in out[java_names[f.group(1)]])))
            elif s:
                fields.append((s.group(3), int(s.group(1)) * SIZES[s.group(2)]))
            elif p:
                fields.append((None, int(p.group(1))))
            else:
                raise AssertionError(f"unparsed layout line: {line!r}")
        out[m.group(3)] = fields
        java_names[m.group(1)] = m.group(3)
    return out


def test_panama_layouts_mirror_the_header_field_by_field():
    src = open(os.path.join(JAVA, "VsrLayouts.java")).read()
    layouts = _parse_layouts(src)
    structs = {"vsr_settings": A.vsr_settings, "vsr_material": A.vsr_material, "vsr_sensor": A.vsr_sensor,
               "vsr_breakable": A.vsr_breakable, "vsr_shape": A.vsr_shape, "vsr_batch_desc": A.vsr_batch_desc, "vsr_result": A.vsr_result}
    assert set(layouts) == set(structs)
    for name, ct in structs.items():
        off = 0
        named = [(f, s) for f, s in layouts[name]]
        cfields = [f[0] for f in ct._fields_]
        assert [f for f, _ in named if f is not None] == cfields, name       # same names, same order
        for f, size in named:
            if f is not None:
                cf = getattr(ct, f)
                assert (cf.offset, cf.size) == (off, size), (name, f, cf.offset, off, cf.size, size)
            off += size
        assert off == C.sizeof(ct), name


def test_java_constants_equal_the_header():
    src = open(os.path.join(JAVA, "VsrLayouts.java")).read()
    consts = {k: int(v) for k, v in re.findall(r"\b([A-Z][A-Z0-9_]+) = (-?\d+)[,;]", src)}
    hdr = open(os.path.join(os.path.dirname(JAVA), "..", "..", "..", "..", "..", "include", "vsr_b200.h")).read()
    macros = {k: int(v.rstrip("u")) for k, v in re.findall(r"#define (VSR_\w+) \(?(-?\d+u?)\)?", hdr)}
    rename = {"SENSOR_AREA_RATIO": "VSR_SENSOR_AREA_RATIO", "OK": "VSR_OK"}
    checked = 0
    for k, v in consts.items():
        h = rename.get(k, k if k.startswith("VSR_") else "VSR_" + k)
        if h in macros:
            assert macros[h] == v, (k, h, v, macros[h])
            checked += 1
    assert checked >= 40
    # every symbol the binding looks up is exported by the header
    bind = open(os.path.join(JAVA, "VsrB200.java")).read()
    for sym in re.findall(r'h\("(\w+)"', bind):
        assert re.search(r"\b" + sym + r"\s*\(", hdr), sym
