"""librogtk_polars_plugin.so: the `_polars_plugin_*` symbols the reference's #[polars_expr] functions generate (reference
src/expressions.rs:695, :770, :880; src/fracture_opt.rs:283), implemented in C++ (rogtk_b200/csrc/polars_plugin.cpp).

Polars is not installable in the build image, so this harness plays Polars: it pickles the kwargs dict exactly as
register_plugin_function does (reference rogtk/__init__.py:138-156), exports pyarrow arrays through the Arrow C Data Interface into
fabricated polars_ffi SeriesExport structs, calls the symbol, and imports the returned Series back with pyarrow."""
import ctypes
import os
import pickle

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "rogtk_b200", "librogtk_polars_plugin.so")


class ArrowSchema(ctypes.Structure):
    pass


class ArrowArray(ctypes.Structure):
    pass


ArrowSchema._fields_ = [("format", ctypes.c_char_p), ("name", ctypes.c_char_p), ("metadata", ctypes.c_char_p), ("flags", ctypes.c_int64),
                        ("n_children", ctypes.c_int64), ("children", ctypes.POINTER(ctypes.POINTER(ArrowSchema))),
                        ("dictionary", ctypes.POINTER(ArrowSchema)), ("release", ctypes.c_void_p), ("private_data", ctypes.c_void_p)]
ArrowArray._fields_ = [("length", ctypes.c_int64), ("null_count", ctypes.c_int64), ("offset", ctypes.c_int64), ("n_buffers", ctypes.c_int64),
                       ("n_children", ctypes.c_int64), ("buffers", ctypes.POINTER(ctypes.c_void_p)),
                       ("children", ctypes.POINTER(ctypes.POINTER(ArrowArray))), ("dictionary", ctypes.POINTER(ArrowArray)),
                       ("release", ctypes.c_void_p), ("private_data", ctypes.c_void_p)]


class SeriesExport(ctypes.Structure):
    _fields_ = [("field", ctypes.POINTER(ArrowSchema)), ("arrays", ctypes.POINTER(ctypes.POINTER(ArrowArray))), ("len", ctypes.c_size_t),
                ("release", ctypes.c_void_p), ("private_data", ctypes.c_void_p)]


RELEASE_SERIES = ctypes.CFUNCTYPE(None, ctypes.POINTER(SeriesExport))


@pytest.fixture(scope="module")
def plugin():
    if not os.path.exists(PLUGIN):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(PLUGIN)
    lib._polars_plugin_get_last_error_message.restype = ctypes.c_char_p
    lib._polars_plugin_get_version.restype = ctypes.c_uint32
    lib.rogtk_plugin_debug_parse_kwargs.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    return lib


def _series(name, chunks):
    """fabricate a polars_ffi SeriesExport from pyarrow arrays (one per chunk); returns (struct, keepalive)"""
    import pyarrow as pa
    field = ArrowSchema()
    pa.field(name, chunks[0].type)._export_to_c(ctypes.addressof(field))
    arrs = [ArrowArray() for _ in chunks]
    for a, c in zip(arrs, chunks):
        c._export_to_c(ctypes.addressof(a))
    ptrs = (ctypes.POINTER(ArrowArray) * len(arrs))(*[ctypes.pointer(a) for a in arrs])
    s = SeriesExport(field=ctypes.pointer(field), arrays=ptrs, len=len(arrs), release=None, private_data=None)
    return s, (field, arrs, ptrs)


def _call(lib, fn, inputs, kwargs):
    import pyarrow as pa
    kw = pickle.dumps(kwargs, protocol=5)
    arr_t = SeriesExport * len(inputs)
    keep = []
    ins = arr_t()
    for i, (name, chunks) in enumerate(inputs):
        s, k = _series(name, chunks)
        ins[i] = s
        keep.append(k)
    ret = SeriesExport()
    f = getattr(lib, "_polars_plugin_" + fn)
    f.argtypes = [ctypes.POINTER(SeriesExport), ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(SeriesExport), ctypes.c_void_p]
    f.restype = None
    f(ins, len(inputs), kw, len(kw), ctypes.byref(ret), None)
    # release what the harness exported (Polars keeps ownership of the inputs; the plugin only borrows them)
    if not ret.release:
        return None, lib._polars_plugin_get_last_error_message().decode()
    field = pa.Field._import_from_c(ctypes.addressof(ret.field.contents))
    out = [pa.Array._import_from_c(ctypes.addressof(ret.arrays[i].contents), field.type) for i in range(ret.len)]
    return (field, out), None


def test_plugin_exports_and_kwargs_reader(plugin):
    """The symbols pyo3-polars would generate exist; the pickle-subset reader decodes what Python's pickle writes for the kwargs
    dicts of the reference's four wrappers (reference rogtk/__init__.py:140-156, :218-234, :271-287, :308-324), protocols 2..5."""
    for fn in ("assemble_sequences_expr", "assemble_sequences_with_anchors_expr", "sweep_assembly_params_expr", "optimize_assembly_expr"):
        assert hasattr(plugin, "_polars_plugin_" + fn) and hasattr(plugin, "_polars_plugin_field_" + fn)
    assert plugin._polars_plugin_get_version() == 1
    kwargs = {"k": 31, "min_coverage": 2, "method": "shortest_path", "start_anchor": "GAGACTGCATGG", "end_anchor": None, "min_length": 70000,
              "export_graphs": False, "only_largest": True, "auto_k": False, "prefix": "sample_é", "big": 2 ** 40, "neg": -3}
    want = "".join(f"{k}={v!r};" for k, v in sorted(kwargs.items()))
    for proto in (2, 3, 4, 5):
        buf = ctypes.create_string_buffer(4096)
        blob = pickle.dumps(kwargs, protocol=proto)
        assert plugin.rogtk_plugin_debug_parse_kwargs(blob, len(blob), buf, 4096) == 0
        assert buf.value.decode() == want, proto
    blob = pickle.dumps([1, 2, 3])
    assert plugin.rogtk_plugin_debug_parse_kwargs(blob, len(blob), ctypes.create_string_buffer(64), 64) == -1
    assert b"kwargs" in plugin._polars_plugin_get_last_error_message()


def test_plugin_field_functions(plugin):
    """output_type functions: String named like the input (src/expressions.rs:690-693), the sweep struct (:869-878), the optimize
    struct (src/fracture_opt.rs:357-367)."""
    import pyarrow as pa
    inp = ArrowSchema()
    pa.field("seq", pa.string())._export_to_c(ctypes.addressof(inp))
    for fn, want in (("assemble_sequences_expr", pa.string_view()),
                     ("sweep_assembly_params_expr", pa.struct([("k", pa.int64()), ("min_coverage", pa.int64()), ("contig_length", pa.int64())])),
                     ("optimize_assembly_expr", pa.struct([("contig", pa.string_view()), ("k", pa.uint32()), ("min_coverage", pa.uint32()),
                                                           ("length", pa.uint32()), ("input_sequences", pa.uint32())]))):
        out = ArrowSchema()
        f = getattr(plugin, "_polars_plugin_field_" + fn)
        f.argtypes = [ctypes.POINTER(ArrowSchema), ctypes.c_size_t, ctypes.POINTER(ArrowSchema), ctypes.c_char_p, ctypes.c_size_t]
        f.restype = None
        f(ctypes.byref(inp), 1, ctypes.byref(out), None, 0)
        got = pa.Field._import_from_c(ctypes.addressof(out))
        assert got.name == "seq" and got.type == want, fn


READS = ["GAGACTGCATGGGCTGGTGGGCGTCCGTCTGC", "GGGCTGGTGGGCGTCCGTCTGCTTTAGTGAGGGT"]
WANT = "GAGACTGCATGGGCTGGTGGGCGTCCGTCTGCTTTAGTGAGGGT"


@pytest.mark.gpu
def test_plugin_assemble_one_group_per_call(plugin):
    """The reference's convention: one call per group with a String column.  All three Arrow string layouts, several chunks, nulls; the
    reference's own fixture (src/fracture.rs:611-679) and its error conditions (src/expressions.rs:700-731)."""
    import pyarrow as pa
    kw = dict(k=13, min_coverage=1, method="shortest_path", start_anchor="GAGACTGCATGG", end_anchor="TTTAGTGAGGGT", min_length=None,
              export_graphs=False, only_largest=False, auto_k=False, prefix=None)
    for typ in (pa.string(), pa.large_string(), pa.string_view()):
        (field, out), err = _call(plugin, "assemble_sequences_expr", [("seq", [pa.array(READS + [None], type=typ)])], kw)
        assert err is None and field.name == "assembled_sequences" and out[0].to_pylist() == [WANT], typ
        (field, out), err = _call(plugin, "assemble_sequences_expr", [("seq", [pa.array(READS[:1], type=typ), pa.array([None, READS[1]], type=typ)])], kw)
        assert out[0].to_pylist() == [WANT]
    (field, out), err = _call(plugin, "assemble_sequences_expr", [("seq", [pa.array(READS)])], dict(kw, method="compression", start_anchor=None, end_anchor=None))
    assert out[0].to_pylist() == [WANT]
    (field, out), err = _call(plugin, "assemble_sequences_expr", [("seq", [pa.array(READS)])], dict(kw, start_anchor="AAAAAAAAAAAAA"))
    assert out[0].to_pylist() == [""]
    res, err = _call(plugin, "assemble_sequences_expr", [("seq", [pa.array(READS)])], dict(kw, method="compression"))
    assert res is None and "should not be provided for compression" in err
    res, err = _call(plugin, "assemble_sequences_expr", [("seq", [pa.array(READS)])], dict(kw, method="greedy"))
    assert res is None and "Invalid assembly method" in err
    res, err = _call(plugin, "assemble_sequences_expr", [("seq", [pa.array(READS)])], {"k": 13, "method": "compression"})
    assert res is None and "missing field `min_coverage`" in err
    res, err = _call(plugin, "assemble_sequences_expr", [("seq", [pa.array([1, 2, 3])])], kw)
    assert res is None and "unsupported input dtype" in err
    # dynamic anchors: first row of inputs[1] / inputs[2] (src/expressions.rs:785-797)
    (field, out), err = _call(plugin, "assemble_sequences_with_anchors_expr",
                              [("seq", [pa.array(READS)]), ("s", [pa.array(["GAGACTGCATGG", "x"])]), ("e", [pa.array(["TTTAGTGAGGGT"], type=pa.string_view())])],
                              dict(kw, start_anchor=None, end_anchor=None))
    assert err is None and out[0].to_pylist() == [WANT]
    res, err = _call(plugin, "assemble_sequences_with_anchors_expr", [("seq", [pa.array(READS)])], kw)
    assert res is None and "requires 3 inputs" in err


@pytest.mark.gpu
def test_plugin_batched_list_input_and_struct_outputs(plugin, oracle):
    """The batched form: a List(String) column, one row per group, same function names and kwargs; sweep and optimize return the
    reference's structs, one row per cell / per group."""
    import pyarrow as pa
    groups = [READS, READS[:1], None, [], READS[::-1]]
    lst = pa.array(groups, type=pa.list_(pa.string_view()))
    kw = dict(k=13, min_coverage=1, method="compression", start_anchor=None, end_anchor=None, min_length=None, export_graphs=False,
              only_largest=False, auto_k=False, prefix=None)
    (field, out), err = _call(plugin, "assemble_sequences_expr", [("seq", [lst])], kw)
    assert err is None and out[0].to_pylist() == [WANT, READS[0], "", "", WANT]
    (field, out), err = _call(plugin, "assemble_sequences_expr", [("seq", [lst.slice(0, 2), lst.slice(2)])], kw)     # two chunks
    assert out[0].to_pylist() == [WANT, READS[0], "", "", WANT]
    skw = dict(k_start=10, k_end=20, k_step=5, cov_start=1, cov_end=3, cov_step=1, method="compression", start_anchor=None, end_anchor=None,
               min_length=None, export_graphs=False, prefix=None, auto_k=False)
    (field, out), err = _call(plugin, "sweep_assembly_params_expr", [("seq", [pa.array(READS)])], skw)
    assert err is None and field.name == "seq"
    rows = out[0].to_pylist()
    assert [(r["k"], r["min_coverage"], r["contig_length"]) for r in rows] == oracle.sweep_py(READS, 10, 20, 5, 1, 3, 1, "compression")
    okw = dict(start_anchor="GAGACTGCAT", end_anchor="TTTAGTGAGGGT", method="shortest_path", min_length=None, export_graphs=False, prefix=None,
               start_k=13, start_min_coverage=1, max_iterations=5, explore_k=False, prioritize_length=False)
    (field, out), err = _call(plugin, "optimize_assembly_expr", [("seq", [pa.array(READS + [None])])], okw)
    assert err is None
    (contig, k, cov, length, nseq), _ = oracle.optimize_py(READS, 13, 1, "GAGACTGCAT", "TTTAGTGAGGGT", "shortest_path", 5, False, False)
    assert out[0].to_pylist() == [{"contig": contig, "k": k, "min_coverage": cov, "length": length, "input_sequences": nseq}]
    res, err = _call(plugin, "optimize_assembly_expr", [("seq", [pa.array(READS)])], dict(okw, start_anchor=None))
    assert res is None and "start_anchor is required" in err
