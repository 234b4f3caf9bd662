// polars_plugin.cpp -- librogtk_polars_plugin.so: the Polars expression-plugin boundary of the reference, implemented directly in C++.
//
// The reference exposes its assembly functions as `#[polars_expr]` functions (reference src/expressions.rs:695, :770, :880;
// src/fracture_opt.rs:283); pyo3-polars expands each into C symbols that Polars dlopens through
// register_plugin_function(plugin_path=..., function_name=...) (reference rogtk/__init__.py:138-156):
//
//     void _polars_plugin_<name>(SeriesExport *inputs, size_t n_inputs, const uint8_t *kwargs, size_t kwargs_len,
//                                SeriesExport *ret, CallerContext *ctx);
//     void _polars_plugin_field_<name>(ArrowSchema *fields, size_t n, ArrowSchema *ret, const uint8_t *kwargs, size_t kwargs_len);
//     const char *_polars_plugin_get_last_error_message(void);
//     uint32_t _polars_plugin_get_version(void);
//
// [UPSTREAM-RECALLED: polars_ffi version_0 / pyo3-polars 0.17-0.2x; neither crate nor Polars is available in the build image, so
// the layout below follows SURVEY.md 8(b) and is exercised by tests/test_plugin_abi.py, which fabricates the SeriesExport structs
// with pyarrow's Arrow C Data Interface export -- not by Polars itself.]
//
// Inputs arrive as Arrow C Data Interface arrays (one per chunk), kwargs as a Python pickle of the kwargs dict (decoded here by a
// pickle-subset reader standing in for serde-pickle), results leave as Arrow arrays whose release callbacks free them.  Two
// input shapes per function, the input dtype decides:
//     String / LargeString / StringView column   one call = one group (the reference's convention, group_by().agg())
//     List(String) column                        one call = one frame, one row per group (the batched form a GPU wants)
// The computation itself is rogtk_gpu_assemble_batch / _sweep_batch / _optimize_batch of librogtk_gpu.so (include/rogtk_gpu.h): no
// CPU fallback -- without a B200 the call fails with the library's error text.
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rogtk_gpu.h"

// ---- Arrow C Data Interface (https://arrow.apache.org/docs/format/CDataInterface.html) ---------------------------------------------
extern "C" {
struct ArrowSchema {
    const char *format;
    const char *name;
    const char *metadata;
    int64_t flags;
    int64_t n_children;
    struct ArrowSchema **children;
    struct ArrowSchema *dictionary;
    void (*release)(struct ArrowSchema *);
    void *private_data;
};
struct ArrowArray {
    int64_t length;
    int64_t null_count;
    int64_t offset;
    int64_t n_buffers;
    int64_t n_children;
    const void **buffers;
    struct ArrowArray **children;
    struct ArrowArray *dictionary;
    void (*release)(struct ArrowArray *);
    void *private_data;
};
// polars_ffi::version_0::SeriesExport
struct SeriesExport {
    ArrowSchema *field;
    ArrowArray **arrays;
    size_t len;
    void (*release)(SeriesExport *);
    void *private_data;
};
struct CallerContext;
}

namespace {

thread_local std::string t_last_error;
struct PluginError {
    std::string msg;
};

// ---- kwargs: a pickle-subset reader (what pickle.dumps(dict of str -> int | bool | str | None) produces, protocols 2-5) ----------------
struct PyVal {
    enum Kind { NONE, BOOL, INT, STR, DICT, MARK } kind = NONE;
    int64_t i = 0;
    std::string s;
    std::vector<std::pair<std::string, PyVal>> d;
};
struct Kwargs {
    std::map<std::string, PyVal> kv;
    bool has(const char *k) const
    {
        auto it = kv.find(k);
        return it != kv.end() && it->second.kind != PyVal::NONE;
    }
    int64_t req_int(const char *k) const
    {
        auto it = kv.find(k);
        if (it == kv.end()) throw PluginError{std::string("missing field `") + k + "`"};
        if (it->second.kind != PyVal::INT && it->second.kind != PyVal::BOOL) throw PluginError{std::string("invalid type for `") + k + "`: expected an integer"};
        if (it->second.i < 0) throw PluginError{std::string("invalid value for `") + k + "`: expected usize"};
        return it->second.i;
    }
    std::string req_str(const char *k) const
    {
        auto it = kv.find(k);
        if (it == kv.end()) throw PluginError{std::string("missing field `") + k + "`"};
        if (it->second.kind != PyVal::STR) throw PluginError{std::string("invalid type for `") + k + "`: expected a string"};
        return it->second.s;
    }
    bool opt_str(const char *k, std::string &out) const
    {
        auto it = kv.find(k);
        if (it == kv.end() || it->second.kind == PyVal::NONE) return false;
        if (it->second.kind != PyVal::STR) throw PluginError{std::string("invalid type for `") + k + "`: expected a string"};
        out = it->second.s;
        return true;
    }
    int64_t opt_int(const char *k, int64_t dflt) const
    {
        auto it = kv.find(k);
        if (it == kv.end() || it->second.kind == PyVal::NONE) return dflt;
        if (it->second.kind != PyVal::INT && it->second.kind != PyVal::BOOL) throw PluginError{std::string("invalid type for `") + k + "`"};
        return it->second.i;
    }
};

Kwargs parse_kwargs(const uint8_t *p, size_t n)
{
    Kwargs out;
    if (!p || n == 0) return out;
    std::vector<PyVal> st;
    std::map<uint32_t, PyVal> memo;
    uint32_t next_memo = 0;
    size_t i = 0;
    auto need = [&](size_t k) {
        if (i + k > n) throw PluginError{"kwargs: truncated pickle"};
    };
    auto rd = [&](int bytes) -> uint64_t {
        need((size_t)bytes);
        uint64_t v = 0;
        for (int b = 0; b < bytes; b++) v |= (uint64_t)p[i + b] << (8 * b);
        i += (size_t)bytes;
        return v;
    };
    while (i < n) {
        const uint8_t op = p[i++];
        switch (op) {
        case 0x80: rd(1); break;                                   // PROTO
        case 0x95: rd(8); break;                                   // FRAME
        case '}': { PyVal v; v.kind = PyVal::DICT; st.push_back(v); break; }   // EMPTY_DICT
        case '(': { PyVal v; v.kind = PyVal::MARK; st.push_back(v); break; }   // MARK
        case 0x94: if (!st.empty()) memo[next_memo++] = st.back(); break;       // MEMOIZE
        case 'q': { const uint32_t k = (uint32_t)rd(1); if (!st.empty()) memo[k] = st.back(); break; }   // BINPUT
        case 'r': { const uint32_t k = (uint32_t)rd(4); if (!st.empty()) memo[k] = st.back(); break; }   // LONG_BINPUT
        case 'h': { const uint32_t k = (uint32_t)rd(1); st.push_back(memo[k]); break; }                  // BINGET
        case 'j': { const uint32_t k = (uint32_t)rd(4); st.push_back(memo[k]); break; }                  // LONG_BINGET
        case 0x8c: case 'X': case 0x8d: {                          // SHORT_BINUNICODE, BINUNICODE, BINUNICODE8
            const uint64_t len = rd(op == 0x8c ? 1 : op == 'X' ? 4 : 8);
            need((size_t)len);
            PyVal v; v.kind = PyVal::STR; v.s.assign((const char *)p + i, (size_t)len);
            i += (size_t)len;
            st.push_back(v);
            break;
        }
        case 'K': { PyVal v; v.kind = PyVal::INT; v.i = (int64_t)rd(1); st.push_back(v); break; }        // BININT1
        case 'M': { PyVal v; v.kind = PyVal::INT; v.i = (int64_t)rd(2); st.push_back(v); break; }        // BININT2
        case 'J': { PyVal v; v.kind = PyVal::INT; v.i = (int64_t)(int32_t)rd(4); st.push_back(v); break; }   // BININT
        case 0x8a: {                                               // LONG1
            const int len = (int)rd(1);
            if (len > 8) throw PluginError{"kwargs: integer too large"};
            uint64_t raw = len ? rd(len) : 0;
            if (len && len < 8 && (raw >> (8 * len - 1)) & 1) raw |= ~0ull << (8 * len);   // sign extension
            PyVal v; v.kind = PyVal::INT; v.i = (int64_t)raw; st.push_back(v);
            break;
        }
        case 0x88: { PyVal v; v.kind = PyVal::BOOL; v.i = 1; st.push_back(v); break; }                   // NEWTRUE
        case 0x89: { PyVal v; v.kind = PyVal::BOOL; v.i = 0; st.push_back(v); break; }                   // NEWFALSE
        case 'N': { PyVal v; st.push_back(v); break; }                                                   // NONE
        case 's': {                                                // SETITEM
            if (st.size() < 3) throw PluginError{"kwargs: bad pickle (SETITEM)"};
            PyVal val = st.back(); st.pop_back();
            PyVal key = st.back(); st.pop_back();
            if (st.back().kind != PyVal::DICT || key.kind != PyVal::STR) throw PluginError{"kwargs: expected a dict of str"};
            st.back().d.emplace_back(key.s, val);
            break;
        }
        case 'u': {                                                // SETITEMS
            size_t m = st.size();
            while (m > 0 && st[m - 1].kind != PyVal::MARK) m--;
            if (m == 0 || m < 2 || st[m - 2].kind != PyVal::DICT || ((st.size() - m) & 1)) throw PluginError{"kwargs: bad pickle (SETITEMS)"};
            for (size_t j = m; j + 1 < st.size(); j += 2) {
                if (st[j].kind != PyVal::STR) throw PluginError{"kwargs: expected a dict of str"};
                st[m - 2].d.emplace_back(st[j].s, st[j + 1]);
            }
            st.resize(m - 1);
            break;
        }
        case '.':                                                  // STOP
            if (st.empty() || st.back().kind != PyVal::DICT) throw PluginError{"kwargs: expected a dict"};
            for (auto &e : st.back().d) out.kv[e.first] = e.second;
            return out;
        default:
            throw PluginError{"kwargs: unsupported pickle opcode " + std::to_string((int)op)};
        }
    }
    throw PluginError{"kwargs: pickle without STOP"};
}

// ---- GPU contexts: one per thread of the caller (Polars runs plugins concurrently from its rayon workers) ---------------------------
struct CtxHolder {
    rogtk_gpu_ctx *ctx = nullptr;
    ~CtxHolder()
    {
        if (ctx) rogtk_gpu_destroy(ctx);
    }
};
rogtk_gpu_ctx *thread_ctx()
{
    thread_local CtxHolder h;
    if (!h.ctx) {
        int err = 0;
        const char *dev = getenv("ROGTK_PLUGIN_DEVICE");
        h.ctx = rogtk_gpu_create(dev ? atoi(dev) : 0, 0, &err);
        if (!h.ctx) throw PluginError{"rogtk_gpu_create failed (error " + std::to_string(err) + "): this build has no CPU fallback, a B200 is required"};
    }
    return h.ctx;
}

// ---- input: a String-like Series (chunks) or a List(String) Series -> one host strcol + group offsets ---------------------------------
struct Column {
    rogtk_strcol col{};
    std::vector<int64_t> group_offsets;      // n_groups + 1
    // storage when the chunks had to be concatenated
    std::vector<uint8_t> data, validity;
    std::vector<int64_t> offsets;
    std::vector<const uint8_t *> bufs;
    std::vector<int64_t> buf_len;
    bool batched = false;                    // List(String) input: one row per group
};

bool is_str_format(const char *f) { return f && (!strcmp(f, "u") || !strcmp(f, "U") || !strcmp(f, "vu") || !strcmp(f, "z") || !strcmp(f, "Z") || !strcmp(f, "vz")); }

bool bit_at(const uint8_t *bm, int64_t i) { return !bm || ((bm[i >> 3] >> (i & 7)) & 1); }

// appends the rows [0, length) of one string array (any of the three layouts) to a LargeUtf8 column under construction
void append_string_array(const char *fmt, const ArrowArray *a, std::vector<uint8_t> &data, std::vector<int64_t> &offsets, std::vector<uint8_t> &valid_bits)
{
    const uint8_t *bm = a->n_buffers > 0 ? (const uint8_t *)a->buffers[0] : nullptr;
    for (int64_t r = 0; r < a->length; r++) {
        const int64_t i = a->offset + r;
        const bool ok = bit_at(bm, i);
        valid_bits.push_back(ok ? 1 : 0);
        if (ok) {
            const uint8_t *src = nullptr;
            int64_t len = 0;
            if (!strcmp(fmt, "u") || !strcmp(fmt, "z")) {
                const int32_t *o = (const int32_t *)a->buffers[1];
                src = (const uint8_t *)a->buffers[2] + o[i];
                len = o[i + 1] - o[i];
            } else if (!strcmp(fmt, "U") || !strcmp(fmt, "Z")) {
                const int64_t *o = (const int64_t *)a->buffers[1];
                src = (const uint8_t *)a->buffers[2] + o[i];
                len = o[i + 1] - o[i];
            } else {
                const uint8_t *v = (const uint8_t *)a->buffers[1] + 16 * i;
                uint32_t l, bi, off;
                memcpy(&l, v, 4);
                len = l;
                if (l <= 12) {
                    src = v + 4;
                } else {
                    memcpy(&bi, v + 8, 4);
                    memcpy(&off, v + 12, 4);
                    if ((int64_t)bi + 2 >= a->n_buffers) throw PluginError{"malformed string view: buffer index out of range"};
                    src = (const uint8_t *)a->buffers[2 + bi] + off;
                }
            }
            data.insert(data.end(), src, src + len);
        }
        offsets.push_back((int64_t)data.size());
    }
}

void finish_concat(Column &c, const std::vector<uint8_t> &valid_bits)
{
    const int64_t n = (int64_t)valid_bits.size();
    bool any_null = false;
    for (uint8_t b : valid_bits) any_null |= !b;
    if (any_null) {
        c.validity.assign((size_t)(n + 7) / 8, 0);
        for (int64_t r = 0; r < n; r++)
            if (valid_bits[(size_t)r]) c.validity[(size_t)r >> 3] |= (uint8_t)(1u << (r & 7));
    }
    c.bufs = {c.data.empty() ? (const uint8_t *)"" : c.data.data()};
    c.buf_len = {(int64_t)c.data.size()};
    c.col.layout = ROGTK_UTF8_I64;
    c.col.n_buffers = 1;
    c.col.n_rows = n;
    c.col.validity = any_null ? c.validity.data() : nullptr;
    c.col.validity_bit_offset = 0;
    c.col.offsets_or_views = c.offsets.data();
    c.col.data = c.bufs.data();
    c.col.data_len = c.buf_len.data();
}

void read_input(const SeriesExport *s, Column &c)
{
    if (!s || !s->field || !s->field->format) throw PluginError{"input series without a schema"};
    const char *fmt = s->field->format;
    if (is_str_format(fmt)) {
        // one group.  A single unsliced chunk of an offsets layout goes to the library as it is (zero copy); everything else is
        // concatenated into LargeUtf8 on the host (views keep the zero-copy route only through the batched entry).
        if (s->len == 1 && s->arrays[0]->offset == 0 && (!strcmp(fmt, "u") || !strcmp(fmt, "U") || !strcmp(fmt, "z") || !strcmp(fmt, "Z"))) {
            const ArrowArray *a = s->arrays[0];
            const bool large = fmt[0] == 'U' || fmt[0] == 'Z';
            c.bufs = {(const uint8_t *)a->buffers[2]};
            const int64_t last = large ? ((const int64_t *)a->buffers[1])[a->length] : (int64_t)((const int32_t *)a->buffers[1])[a->length];
            c.buf_len = {last};
            c.col.layout = large ? ROGTK_UTF8_I64 : ROGTK_UTF8_I32;
            c.col.n_buffers = 1;
            c.col.n_rows = a->length;
            c.col.validity = a->null_count != 0 ? (const uint8_t *)a->buffers[0] : nullptr;
            c.col.validity_bit_offset = 0;
            c.col.offsets_or_views = a->buffers[1];
            c.col.data = c.bufs.data();
            c.col.data_len = c.buf_len.data();
        } else {
            std::vector<uint8_t> vb;
            c.offsets.push_back(0);
            for (size_t k = 0; k < s->len; k++) append_string_array(fmt, s->arrays[k], c.data, c.offsets, vb);
            finish_concat(c, vb);
        }
        c.group_offsets = {0, c.col.n_rows};
        return;
    }
    if (!strcmp(fmt, "+l") || !strcmp(fmt, "+L")) {
        if (s->field->n_children != 1 || !is_str_format(s->field->children[0]->format)) throw PluginError{"expected a List(String) column"};
        const char *cf = s->field->children[0]->format;
        const bool large = fmt[1] == 'L';
        std::vector<uint8_t> vb;
        c.offsets.push_back(0);
        c.group_offsets.push_back(0);
        for (size_t k = 0; k < s->len; k++) {
            const ArrowArray *a = s->arrays[k];
            const uint8_t *bm = a->n_buffers > 0 ? (const uint8_t *)a->buffers[0] : nullptr;
            const ArrowArray *child = a->children[0];
            for (int64_t r = 0; r < a->length; r++) {
                const int64_t i = a->offset + r;
                int64_t lo = large ? ((const int64_t *)a->buffers[1])[i] : ((const int32_t *)a->buffers[1])[i];
                int64_t hi = large ? ((const int64_t *)a->buffers[1])[i + 1] : ((const int32_t *)a->buffers[1])[i + 1];
                if (!bit_at(bm, i)) hi = lo;   // a null list is an empty group
                ArrowArray view = *child;
                view.offset = child->offset + lo;
                view.length = hi - lo;
                append_string_array(cf, &view, c.data, c.offsets, vb);
                c.group_offsets.push_back((int64_t)vb.size());
            }
        }
        finish_concat(c, vb);
        c.batched = true;
        return;
    }
    throw PluginError{std::string("unsupported input dtype (Arrow format `") + fmt + "`): expected String or List(String)"};
}

// first non-null value of a String series (assemble_sequences_with_anchors_expr: `.get(0)`, reference src/expressions.rs:785-797)
bool first_string(const SeriesExport *s, std::string &out)
{
    if (!s || !s->field || !is_str_format(s->field->format)) throw PluginError{"anchor input is not a String column"};
    for (size_t k = 0; k < s->len; k++) {
        const ArrowArray *a = s->arrays[k];
        if (a->length == 0) continue;
        std::vector<uint8_t> data, vb;
        std::vector<int64_t> off{0};
        ArrowArray one = *a;
        one.length = 1;
        append_string_array(s->field->format, &one, data, off, vb);
        if (!vb[0]) return false;
        out.assign(data.begin(), data.end());
        return true;
    }
    return false;
}

// ---- output: Arrow arrays whose release callbacks own the memory ----------------------------------------------------------------------
struct OwnedArray {
    std::vector<std::vector<uint8_t>> bufs;
    std::vector<const void *> ptrs;
    std::vector<ArrowArray *> children;
};
void release_array(ArrowArray *a)
{
    if (!a || !a->release) return;
    OwnedArray *o = (OwnedArray *)a->private_data;
    for (ArrowArray *ch : o->children) {
        if (ch->release) ch->release(ch);
        delete ch;
    }
    delete o;
    a->release = nullptr;
}
struct OwnedSchema {
    std::string format, name;
    std::vector<ArrowSchema *> children;
};
void release_schema(ArrowSchema *s)
{
    if (!s || !s->release) return;
    OwnedSchema *o = (OwnedSchema *)s->private_data;
    for (ArrowSchema *ch : o->children) {
        if (ch->release) ch->release(ch);
        delete ch;
    }
    delete o;
    s->release = nullptr;
}
void make_schema(ArrowSchema *s, const char *format, const char *name, std::vector<ArrowSchema *> children = {})
{
    OwnedSchema *o = new OwnedSchema{format, name ? name : "", children};
    memset(s, 0, sizeof(*s));
    s->format = o->format.c_str();
    s->name = o->name.c_str();
    s->flags = 2;   // ARROW_FLAG_NULLABLE
    s->n_children = (int64_t)o->children.size();
    s->children = o->children.empty() ? nullptr : o->children.data();
    s->release = release_schema;
    s->private_data = o;
}
ArrowArray *make_primitive(const void *data, size_t bytes, int64_t n)
{
    OwnedArray *o = new OwnedArray();
    o->bufs.emplace_back((const uint8_t *)data, (const uint8_t *)data + bytes);
    o->ptrs = {nullptr, o->bufs[0].data()};
    ArrowArray *a = new ArrowArray();
    memset(a, 0, sizeof(*a));
    a->length = n;
    a->n_buffers = 2;
    a->buffers = o->ptrs.data();
    a->release = release_array;
    a->private_data = o;
    return a;
}
ArrowArray *make_large_utf8(const int64_t *offsets, const uint8_t *data, int64_t n)
{
    OwnedArray *o = new OwnedArray();
    o->bufs.emplace_back((const uint8_t *)offsets, (const uint8_t *)(offsets + n + 1));
    o->bufs.emplace_back(data, data + offsets[n]);
    if (o->bufs[1].empty()) o->bufs[1].push_back(0);
    o->ptrs = {nullptr, o->bufs[0].data(), o->bufs[1].data()};
    ArrowArray *a = new ArrowArray();
    memset(a, 0, sizeof(*a));
    a->length = n;
    a->n_buffers = 3;
    a->buffers = o->ptrs.data();
    a->release = release_array;
    a->private_data = o;
    return a;
}
ArrowArray *make_struct(std::vector<ArrowArray *> children, int64_t n)
{
    OwnedArray *o = new OwnedArray();
    o->children = children;
    o->ptrs = {nullptr};
    ArrowArray *a = new ArrowArray();
    memset(a, 0, sizeof(*a));
    a->length = n;
    a->n_buffers = 1;
    a->buffers = o->ptrs.data();
    a->n_children = (int64_t)o->children.size();
    a->children = o->children.data();
    a->release = release_array;
    a->private_data = o;
    return a;
}
struct SeriesOwner {
    ArrowSchema field;
    ArrowArray *arr;
    ArrowArray *slot[1];
};
void release_series(SeriesExport *s)
{
    if (!s || !s->release) return;
    SeriesOwner *o = (SeriesOwner *)s->private_data;
    if (o->field.release) o->field.release(&o->field);
    if (o->arr) {
        if (o->arr->release) o->arr->release(o->arr);
        delete o->arr;
    }
    delete o;
    s->release = nullptr;
    s->private_data = nullptr;
}
void export_series(SeriesExport *ret, SeriesOwner *o)
{
    o->slot[0] = o->arr;
    ret->field = &o->field;
    ret->arrays = o->slot;
    ret->len = 1;
    ret->release = release_series;
    ret->private_data = o;
}

int method_of(const std::string &m, const char *bad_msg)
{
    if (m == "compression") return ROGTK_METHOD_COMPRESSION;
    if (m == "shortest_path") return ROGTK_METHOD_SHORTEST_PATH;
    if (m == "shortest_path_auto") return ROGTK_METHOD_SHORTEST_PATH_AUTO;
    throw PluginError{bad_msg};
}

template <typename F>
void guarded(SeriesExport *ret, F &&fn)
{
    if (ret) memset(ret, 0, sizeof(*ret));
    try {
        fn();
    } catch (const PluginError &e) {
        t_last_error = e.msg;
        if (ret) memset(ret, 0, sizeof(*ret));
    } catch (const std::exception &e) {
        t_last_error = std::string("the plugin panicked: ") + e.what();
        if (ret) memset(ret, 0, sizeof(*ret));
    } catch (...) {
        t_last_error = "the plugin panicked";
        if (ret) memset(ret, 0, sizeof(*ret));
    }
}

void assemble_common(SeriesExport *inputs, size_t n_inputs, const uint8_t *kw, size_t kwlen, SeriesExport *ret, bool dynamic_anchors)
{
    const Kwargs k = parse_kwargs(kw, kwlen);
    rogtk_assembly_params p{};
    p.k = k.req_int("k");
    p.min_coverage = (uint32_t)std::min<int64_t>(k.req_int("min_coverage"), 0xffffffffll);
    const std::string method = k.req_str("method");
    std::string sa, ea;
    bool has_s = k.opt_str("start_anchor", sa), has_e = k.opt_str("end_anchor", ea);
    if (dynamic_anchors) {
        // reference src/expressions.rs:776-829
        if (n_inputs < 3) throw PluginError{"assemble_sequences_with_anchors requires 3 inputs: sequences, start_anchor, end_anchor"};
        if (!first_string(&inputs[1], sa)) throw PluginError{"start_anchor column is empty"};
        if (!first_string(&inputs[2], ea)) throw PluginError{"end_anchor column is empty"};
        if (method == "compression") throw PluginError{"compression method is not supported with dynamic anchors; use shortest_path"};
        if (method == "shortest_path_auto") throw PluginError{"shortest_path_auto method is not supported with dynamic anchors; use shortest_path"};
        if (method != "shortest_path") throw PluginError{"Invalid assembly method for dynamic anchors. Must be 'shortest_path'"};
        has_s = has_e = true;
    }
    p.method = method_of(method, "Invalid assembly method. Must be 'compression', 'shortest_path', or 'shortest_path_auto'");
    p.start_anchor = has_s ? sa.c_str() : nullptr;
    p.end_anchor = has_e ? ea.c_str() : nullptr;
    p.min_length = k.opt_int("min_length", 0);
    p.only_largest = 1;   // Some(true), hard-coded by both expressions (reference src/expressions.rs:748, :835)
    p.auto_k = k.opt_int("auto_k", 0) != 0;
    if (n_inputs < 1) throw PluginError{"no input series"};
    Column c;
    read_input(&inputs[0], c);
    rogtk_gpu_ctx *ctx = thread_ctx();
    rogtk_string_batch out;
    const int rc = rogtk_gpu_assemble_batch(ctx, &c.col, c.group_offsets.data(), (int64_t)c.group_offsets.size() - 1, &p, 0, &out);
    if (rc != ROGTK_OK) {
        const std::string msg = rogtk_gpu_last_error(ctx);
        throw PluginError{rc == ROGTK_ERR_INVALID_ARG ? msg : "Assembly failed: " + msg};
    }
    SeriesOwner *o = new SeriesOwner();
    make_schema(&o->field, "U", "assembled_sequences");
    o->arr = make_large_utf8(out.offsets, out.data, out.n_rows);
    rogtk_gpu_free_strings(ctx, &out);
    export_series(ret, o);
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

const char *_polars_plugin_get_last_error_message(void) { return t_last_error.c_str(); }
// polars_ffi version (major 0, minor 1): (major << 16) | minor
uint32_t _polars_plugin_get_version(void) { return (0u << 16) | 1u; }

// ---- assemble_sequences_expr (reference src/expressions.rs:695-762) ------------------------------------------------------------------
void _polars_plugin_assemble_sequences_expr(SeriesExport *inputs, size_t n_inputs, const uint8_t *kwargs, size_t kwargs_len, SeriesExport *ret,
                                            CallerContext *)
{
    guarded(ret, [&]() { assemble_common(inputs, n_inputs, kwargs, kwargs_len, ret, false); });
}
// output_string_type (reference src/expressions.rs:690-693): String, named like the first input
void _polars_plugin_field_assemble_sequences_expr(ArrowSchema *fields, size_t n, ArrowSchema *ret, const uint8_t *, size_t)
{
    make_schema(ret, "vu", n > 0 && fields[0].name ? fields[0].name : "");
}

// ---- assemble_sequences_with_anchors_expr (reference src/expressions.rs:770-849) -----------------------------------------------------
void _polars_plugin_assemble_sequences_with_anchors_expr(SeriesExport *inputs, size_t n_inputs, const uint8_t *kwargs, size_t kwargs_len,
                                                         SeriesExport *ret, CallerContext *)
{
    guarded(ret, [&]() { assemble_common(inputs, n_inputs, kwargs, kwargs_len, ret, true); });
}
void _polars_plugin_field_assemble_sequences_with_anchors_expr(ArrowSchema *fields, size_t n, ArrowSchema *ret, const uint8_t *, size_t)
{
    make_schema(ret, "vu", n > 0 && fields[0].name ? fields[0].name : "");
}

// ---- sweep_assembly_params_expr (reference src/expressions.rs:880-955): Struct{k, min_coverage, contig_length: Int64} ------------------
void _polars_plugin_sweep_assembly_params_expr(SeriesExport *inputs, size_t n_inputs, const uint8_t *kwargs, size_t kwargs_len, SeriesExport *ret,
                                               CallerContext *)
{
    guarded(ret, [&]() {
        const Kwargs k = parse_kwargs(kwargs, kwargs_len);
        rogtk_sweep_params p{};
        p.k_start = k.req_int("k_start"); p.k_end = k.req_int("k_end"); p.k_step = k.req_int("k_step");
        p.cov_start = k.req_int("cov_start"); p.cov_end = k.req_int("cov_end"); p.cov_step = k.req_int("cov_step");
        const std::string method = k.req_str("method");
        std::string sa, ea;
        const bool has_s = k.opt_str("start_anchor", sa), has_e = k.opt_str("end_anchor", ea);
        p.method = method_of(method, ("Invalid assembly method: Unknown assembly method: " + method).c_str());
        p.start_anchor = has_s ? sa.c_str() : nullptr;
        p.end_anchor = has_e ? ea.c_str() : nullptr;
        if (n_inputs < 1) throw PluginError{"no input series"};
        Column c;
        read_input(&inputs[0], c);
        rogtk_gpu_ctx *ctx = thread_ctx();
        rogtk_sweep_result out;
        const int rc = rogtk_gpu_sweep_batch(ctx, &c.col, c.group_offsets.data(), (int64_t)c.group_offsets.size() - 1, &p, 0, &out);
        if (rc != ROGTK_OK) {
            const std::string msg = rogtk_gpu_last_error(ctx);
            throw PluginError{rc == ROGTK_ERR_INVALID_ARG ? "Invalid assembly method: " + msg : msg};
        }
        const int64_t n = out.n_groups * out.n_cells;
        SeriesOwner *o = new SeriesOwner();
        ArrowSchema *f[3];
        const char *names[3] = {"k", "min_coverage", "contig_length"};
        for (int i = 0; i < 3; i++) {
            f[i] = new ArrowSchema();
            make_schema(f[i], "l", names[i]);
        }
        make_schema(&o->field, "+s", inputs[0].field && inputs[0].field->name ? inputs[0].field->name : "", {f[0], f[1], f[2]});
        o->arr = make_struct({make_primitive(out.k, (size_t)n * 8, n), make_primitive(out.min_coverage, (size_t)n * 8, n),
                              make_primitive(out.contig_length, (size_t)n * 8, n)}, n);
        rogtk_gpu_free_sweep(ctx, &out);
        export_series(ret, o);
    });
}
void _polars_plugin_field_sweep_assembly_params_expr(ArrowSchema *fields, size_t n, ArrowSchema *ret, const uint8_t *, size_t)
{
    ArrowSchema *f[3];
    const char *names[3] = {"k", "min_coverage", "contig_length"};
    for (int i = 0; i < 3; i++) {
        f[i] = new ArrowSchema();
        make_schema(f[i], "l", names[i]);
    }
    make_schema(ret, "+s", n > 0 && fields[0].name ? fields[0].name : "", {f[0], f[1], f[2]});
}

// ---- optimize_assembly_expr (reference src/fracture_opt.rs:283-367): Struct{contig: String, k, min_coverage, length, input_sequences: UInt32}
void _polars_plugin_optimize_assembly_expr(SeriesExport *inputs, size_t n_inputs, const uint8_t *kwargs, size_t kwargs_len, SeriesExport *ret,
                                           CallerContext *)
{
    guarded(ret, [&]() {
        const Kwargs k = parse_kwargs(kwargs, kwargs_len);
        rogtk_optimize_params p{};
        const std::string method = k.req_str("method");
        p.start_k = k.req_int("start_k");
        p.start_min_coverage = k.req_int("start_min_coverage");
        p.max_iterations = k.opt_int("max_iterations", 0);
        p.explore_k = k.opt_int("explore_k", 0) != 0;
        p.prioritize_length = k.opt_int("prioritize_length", 0) != 0;
        std::string sa, ea;
        if (!k.opt_str("start_anchor", sa)) throw PluginError{"start_anchor is required"};
        if (!k.opt_str("end_anchor", ea)) throw PluginError{"end_anchor is required"};
        p.method = method_of(method, ("Unknown assembly method: " + method).c_str());
        p.start_anchor = sa.c_str();
        p.end_anchor = ea.c_str();
        if (n_inputs < 1) throw PluginError{"no input series"};
        Column c;
        read_input(&inputs[0], c);
        rogtk_gpu_ctx *ctx = thread_ctx();
        rogtk_optimize_result out;
        const int rc = rogtk_gpu_optimize_batch(ctx, &c.col, c.group_offsets.data(), (int64_t)c.group_offsets.size() - 1, &p, 0, &out);
        if (rc != ROGTK_OK) throw PluginError{rogtk_gpu_last_error(ctx)};
        const int64_t n = out.contig.n_rows;
        SeriesOwner *o = new SeriesOwner();
        ArrowSchema *f[5];
        const char *names[5] = {"contig", "k", "min_coverage", "length", "input_sequences"};
        for (int i = 0; i < 5; i++) {
            f[i] = new ArrowSchema();
            make_schema(f[i], i == 0 ? "U" : "I", names[i]);
        }
        make_schema(&o->field, "+s", inputs[0].field && inputs[0].field->name ? inputs[0].field->name : "", {f[0], f[1], f[2], f[3], f[4]});
        o->arr = make_struct({make_large_utf8(out.contig.offsets, out.contig.data, n), make_primitive(out.k, (size_t)n * 4, n),
                              make_primitive(out.min_coverage, (size_t)n * 4, n), make_primitive(out.length, (size_t)n * 4, n),
                              make_primitive(out.input_sequences, (size_t)n * 4, n)}, n);
        rogtk_gpu_free_optimize(ctx, &out);
        export_series(ret, o);
    });
}
void _polars_plugin_field_optimize_assembly_expr(ArrowSchema *fields, size_t n, ArrowSchema *ret, const uint8_t *, size_t)
{
    ArrowSchema *f[5];
    const char *names[5] = {"contig", "k", "min_coverage", "length", "input_sequences"};
    for (int i = 0; i < 5; i++) {
        f[i] = new ArrowSchema();
        make_schema(f[i], i == 0 ? "vu" : "I", names[i]);
    }
    make_schema(ret, "+s", n > 0 && fields[0].name ? fields[0].name : "", {f[0], f[1], f[2], f[3], f[4]});
}

// test hook: decode a kwargs pickle and report "key=value;..." (tests/test_plugin_abi.py checks the reader against pickle.dumps without a GPU)
int rogtk_plugin_debug_parse_kwargs(const uint8_t *kwargs, size_t kwargs_len, char *out, size_t cap)
{
    try {
        const Kwargs k = parse_kwargs(kwargs, kwargs_len);
        std::string s;
        for (auto &e : k.kv) {
            s += e.first + "=";
            switch (e.second.kind) {
            case PyVal::NONE: s += "None"; break;
            case PyVal::BOOL: s += e.second.i ? "True" : "False"; break;
            case PyVal::INT: s += std::to_string(e.second.i); break;
            case PyVal::STR: s += "'" + e.second.s + "'"; break;
            default: s += "?"; break;
            }
            s += ";";
        }
        if (s.size() + 1 > cap) return -2;
        memcpy(out, s.c_str(), s.size() + 1);
        return 0;
    } catch (const PluginError &e) {
        t_last_error = e.msg;
        return -1;
    }
}

}  // extern "C"
#pragma GCC visibility pop
