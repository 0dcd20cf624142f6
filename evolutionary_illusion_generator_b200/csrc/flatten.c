/* _flatten: the genome flattener of genome.py (`flatten_genome`) as a CPython extension - host glue of SURVEY.md §8(f) row 4.
 *
 * flatten(genome, input_keys, output_keys, n_outputs) -> (program bytes, n_slots) | None
 *
 * Builds the flat CPPN program the render kernel interprets (csrc/render.cuh) with exactly the graph the reference builds
 * in `create_cppn` (/root/reference/pytorch_neat/pytorch_neat/cppn.py:168-235): `required_for_output` over all connection
 * keys, disabled connections and connections leaving an output node dropped, children in `genome.connections` insertion
 * order, a node without children = the constant `bias` (no activation, cppn.py:79-80).  Sub-graphs made only of constants
 * are float32 in the reference (`torch.full` is float32, python float * float32 tensor stays float32) and are folded here
 * with float32 arithmetic; where the fold needs a transcendental activation (torch's vectorised float32 sin / exp / tanh /
 * sigmoid, not reproducible bit for bit in C) the function returns None and genome.py falls back to its own flattener,
 * which folds with torch itself.  tests/test_host_logic.py compares the bytes of both flatteners on thousands of genomes.
 *
 * Byte layout = FlatProgram._encode in genome.py.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BLOB_MAGIC 0x45494742
#define OUT_F32_CONST (1 << 30)
enum { SLOT_X = 0, SLOT_Y = 1, SLOT_ONE = 2, SLOT_NODE0 = 3 };
enum { ACT_SIGMOID = 0, ACT_TANH = 1, ACT_ABS = 2, ACT_GAUSS = 3, ACT_IDENTITY = 4, ACT_SIN = 5, ACT_RELU = 6 };
enum { MEMO_NONE = 0, MEMO_SLOT = 1, MEMO_CONST = 2 };
enum { ST_OK = 0, ST_PYTHON = 1, ST_ERROR = 2 };   /* ST_PYTHON: let the Python flattener handle (or report) this genome */

typedef struct { int src; double w; int next; } Edge;
typedef struct {
    long key;
    PyObject* keyobj;      /* borrowed */
    int in_head, in_tail;  /* incoming edge list (insertion order), -1 = empty */
    char has_incoming, needed, is_in, is_out, seen, memo;
    int slot;
    float cval;
} KeyInfo;
typedef struct { int act, agg, t0, nt; double bias, resp; } NodeRec;
typedef struct { double w; int slot; } TermRec;

typedef struct {
    KeyInfo* k; int nk, capk;
    int* table; int tsize;       /* open addressing: key -> index into k */
    Edge* e; int ne, cape;
    NodeRec* nodes; int nn, capn;
    TermRec* terms; int nt, capt;
    PyObject* nodes_dict;
    int status;
} Ctx;

static int grow(void** p, int* cap, int need, size_t sz) {
    if (need <= *cap) return 0;
    int nc = *cap ? *cap * 2 : 64;
    while (nc < need) nc *= 2;
    void* q = realloc(*p, (size_t)nc * sz);
    if (!q) return -1;
    *p = q; *cap = nc;
    return 0;
}

static unsigned hash_long(long v) { uint64_t x = (uint64_t)v * 0x9E3779B97F4A7C15ull; return (unsigned)(x >> 40); }

static int rehash(Ctx* c) {
    int ts = c->tsize ? c->tsize * 2 : 256;
    int* t = (int*)malloc(sizeof(int) * ts);
    if (!t) return -1;
    for (int i = 0; i < ts; ++i) t[i] = -1;
    for (int i = 0; i < c->nk; ++i) {
        unsigned h = hash_long(c->k[i].key) & (ts - 1);
        while (t[h] >= 0) h = (h + 1) & (ts - 1);
        t[h] = i;
    }
    free(c->table);
    c->table = t; c->tsize = ts;
    return 0;
}

/* index of `key`, created on first sight */
static int key_id(Ctx* c, long key, PyObject* obj) {
    if (c->nk * 2 >= c->tsize && rehash(c)) return -1;
    unsigned h = hash_long(key) & (c->tsize - 1);
    while (c->table[h] >= 0) {
        if (c->k[c->table[h]].key == key) return c->table[h];
        h = (h + 1) & (c->tsize - 1);
    }
    if (grow((void**)&c->k, &c->capk, c->nk + 1, sizeof(KeyInfo))) return -1;
    KeyInfo* ki = &c->k[c->nk];
    memset(ki, 0, sizeof *ki);
    ki->key = key; ki->keyobj = obj; ki->in_head = ki->in_tail = -1;
    c->table[h] = c->nk;
    return c->nk++;
}

/* attribute names, interned once (PyObject_GetAttrString builds a new str per call) */
static PyObject *s_bias, *s_response, *s_aggregation, *s_activation, *s_enabled, *s_key, *s_weight, *s_connections, *s_nodes;

static int attr_double(PyObject* o, PyObject* name, double* out) {
    PyObject* v = PyObject_GetAttr(o, name);
    if (!v) return -1;
    *out = PyFloat_AsDouble(v);
    Py_DECREF(v);
    return (*out == -1.0 && PyErr_Occurred()) ? -1 : 0;
}

static int act_id(PyObject* s) {
    static const char* names[7] = {"sigmoid", "tanh", "abs", "gauss", "identity", "sin", "relu"};
    if (!PyUnicode_Check(s)) return -1;
    for (int i = 0; i < 7; ++i) if (PyUnicode_CompareWithASCIIString(s, names[i]) == 0) return i;
    return -1;
}

static int emit(Ctx* c, int act, int agg, const TermRec* t, int nt, double bias, double resp) {
    if (grow((void**)&c->terms, &c->capt, c->nt + nt, sizeof(TermRec)) || grow((void**)&c->nodes, &c->capn, c->nn + 1, sizeof(NodeRec))) {
        c->status = ST_ERROR; PyErr_NoMemory(); return -1;
    }
    NodeRec* n = &c->nodes[c->nn];
    n->act = act; n->agg = agg; n->t0 = c->nt; n->nt = nt; n->bias = bias; n->resp = resp;
    memcpy(c->terms + c->nt, t, sizeof(TermRec) * nt);
    c->nt += nt;
    return SLOT_NODE0 + c->nn++;
}

/* memoised evaluation of node `id`: MEMO_SLOT (k->slot) or MEMO_CONST (k->cval, float32) */
static int visit(Ctx* c, int id, int depth) {
    KeyInfo* ki = &c->k[id];
    if (ki->memo) return 0;
    if (depth > 2000 || !ki->has_incoming) { c->status = ST_PYTHON; return -1; }   /* cycle / unknown key: Python raises */
    PyObject* gene = PyDict_GetItemWithError(c->nodes_dict, ki->keyobj);            /* borrowed */
    if (!gene) { c->status = PyErr_Occurred() ? ST_ERROR : ST_PYTHON; return -1; }
    double bias, resp;
    if (attr_double(gene, s_bias, &bias) || attr_double(gene, s_response, &resp)) { c->status = ST_ERROR; return -1; }
    if (ki->in_head < 0) { ki->memo = MEMO_CONST; ki->cval = (float)bias; return 0; }   /* torch.full(shape, bias): float32 */
    PyObject* so = PyObject_GetAttr(gene, s_aggregation);
    if (!so) { c->status = ST_ERROR; return -1; }
    int agg = -1;
    if (PyUnicode_Check(so)) agg = PyUnicode_CompareWithASCIIString(so, "sum") == 0 ? 0 : PyUnicode_CompareWithASCIIString(so, "prod") == 0 ? 1 : -1;
    Py_DECREF(so);
    so = PyObject_GetAttr(gene, s_activation);
    if (!so) { c->status = ST_ERROR; return -1; }
    const int act = act_id(so);
    Py_DECREF(so);
    if (agg < 0 || act < 0) { c->status = ST_PYTHON; return -1; }                   /* Python raises KeyError with the name */
    int n_src = 0, n_var = 0;
    for (int e = ki->in_head; e >= 0; e = c->e[e].next) {
        if (visit(c, c->e[e].src, depth + 1)) return -1;
        ++n_src;
        if (c->k[c->e[e].src].memo == MEMO_SLOT) ++n_var;
    }
    ki = &c->k[id];   /* the table may have moved */
    TermRec stack_terms[64];
    TermRec* terms = n_src + 1 <= 64 ? stack_terms : (TermRec*)malloc(sizeof(TermRec) * (n_src + 1));
    if (!terms) { c->status = ST_ERROR; PyErr_NoMemory(); return -1; }
    int nt = 0, rc = 0;
    if (n_var == n_src) {                                    /* the common case: every source depends on the inputs */
        for (int e = ki->in_head; e >= 0; e = c->e[e].next) { terms[nt].w = c->e[e].w; terms[nt].slot = c->k[c->e[e].src].slot; ++nt; }
        const int slot = emit(c, act, agg, terms, nt, bias, resp);
        if (slot < 0) rc = -1; else { ki = &c->k[id]; ki->memo = MEMO_SLOT; ki->slot = slot; }
    } else if (n_var == 0) {                                 /* constant sub-graph: float32 arithmetic like torch */
        float acc = 0.f;
        int first = 1;
        for (int e = ki->in_head; e >= 0; e = c->e[e].next) {
            const float term = (float)c->e[e].w * c->k[c->e[e].src].cval;
            acc = first ? term : (agg == 0 ? acc + term : acc * term);
            first = 0;
        }
        const float pre = (float)resp * acc + (float)bias;
        if (act == ACT_IDENTITY) { ki->memo = MEMO_CONST; ki->cval = pre; }
        else if (act == ACT_ABS) { ki->memo = MEMO_CONST; ki->cval = fabsf(pre); }
        else if (act == ACT_RELU) { ki->memo = MEMO_CONST; ki->cval = pre > 0.f ? pre : (pre != pre ? pre : 0.f); }
        else { c->status = ST_PYTHON; rc = -1; }             /* torch's float32 sin / exp / tanh / sigmoid */
    } else {                                                 /* constants before the first variable source fold into a prefix */
        float prefix = 0.f;
        int have_prefix = 0, seen_var = 0;
        for (int e = ki->in_head; e >= 0; e = c->e[e].next) {
            const KeyInfo* s = &c->k[c->e[e].src];
            if (s->memo == MEMO_CONST) {
                const float term = (float)c->e[e].w * s->cval;
                if (!seen_var) { prefix = have_prefix ? (agg == 0 ? prefix + term : prefix * term) : term; have_prefix = 1; }
                else { terms[nt].w = (double)term; terms[nt].slot = SLOT_ONE; ++nt; }
            } else {
                if (!seen_var && have_prefix) { terms[nt].w = (double)prefix; terms[nt].slot = SLOT_ONE; ++nt; }
                seen_var = 1;
                terms[nt].w = c->e[e].w; terms[nt].slot = s->slot; ++nt;
            }
        }
        const int slot = emit(c, act, agg, terms, nt, bias, resp);
        if (slot < 0) rc = -1; else { ki = &c->k[id]; ki->memo = MEMO_SLOT; ki->slot = slot; }
    }
    if (terms != stack_terms) free(terms);
    return rc;
}

static void ctx_free(Ctx* c) { free(c->k); free(c->table); free(c->e); free(c->nodes); free(c->terms); }

static int long_of(PyObject* o, long* out) {
    if (!PyLong_Check(o)) return -1;
    *out = PyLong_AsLong(o);
    return (*out == -1 && PyErr_Occurred()) ? -1 : 0;
}

static PyObject* py_flatten(PyObject* self, PyObject* args) {
    PyObject *genome, *in_keys, *out_keys, *n_out_obj;
    if (!PyArg_ParseTuple(args, "OOOO", &genome, &in_keys, &out_keys, &n_out_obj)) return NULL;
    PyObject* conns = PyObject_GetAttr(genome, s_connections);
    PyObject* nodes = conns ? PyObject_GetAttr(genome, s_nodes) : NULL;
    PyObject* result = NULL;
    Ctx c;
    memset(&c, 0, sizeof c);
    int* frontier = NULL; int* layer = NULL;
    if (!conns || !nodes) goto done;
    if (!PyDict_Check(conns) || !PyDict_Check(nodes) || !PyList_Check(in_keys) || !PyList_Check(out_keys) || PyList_GET_SIZE(in_keys) != 2) {
        c.status = ST_PYTHON; goto done;
    }
    c.nodes_dict = nodes;
    if (rehash(&c)) { PyErr_NoMemory(); c.status = ST_ERROR; goto done; }
    const Py_ssize_t n_out_total = PyList_GET_SIZE(out_keys);
    Py_ssize_t n_used = n_out_total;
    if (n_out_obj != Py_None) {
        n_used = PyLong_AsSsize_t(n_out_obj);
        if (n_used == -1 && PyErr_Occurred()) { c.status = ST_ERROR; goto done; }
        if (n_used > n_out_total) n_used = n_out_total;
        if (n_used < 0) { c.status = ST_PYTHON; goto done; }
    }
    long kv;
    for (int i = 0; i < 2; ++i) {
        PyObject* o = PyList_GET_ITEM(in_keys, i);
        if (long_of(o, &kv)) { PyErr_Clear(); c.status = ST_PYTHON; goto done; }
        const int id = key_id(&c, kv, o);
        if (id < 0) { PyErr_NoMemory(); c.status = ST_ERROR; goto done; }
        c.k[id].is_in = 1; c.k[id].memo = MEMO_SLOT; c.k[id].slot = i == 0 ? SLOT_X : SLOT_Y;
    }
    for (Py_ssize_t i = 0; i < n_out_total; ++i) {
        PyObject* o = PyList_GET_ITEM(out_keys, i);
        if (long_of(o, &kv)) { PyErr_Clear(); c.status = ST_PYTHON; goto done; }
        const int id = key_id(&c, kv, o);
        if (id < 0) { PyErr_NoMemory(); c.status = ST_ERROR; goto done; }
        if (c.k[id].is_in) { c.status = ST_PYTHON; goto done; }
        c.k[id].is_out = 1; c.k[id].has_incoming = 1; c.k[id].needed = 1; c.k[id].seen = 1;
    }
    /* pass 1 over the connection dict: predecessor lists by dict KEY (what neat.graphs.required_for_output walks) */
    const Py_ssize_t n_conn = PyDict_GET_SIZE(conns);
    int* pa = (int*)malloc(sizeof(int) * (size_t)(2 * n_conn + 2));
    if (!pa) { PyErr_NoMemory(); c.status = ST_ERROR; goto done; }
    int* pb = pa + n_conn + 1;
    {
        Py_ssize_t pos = 0, i = 0;
        PyObject *dk, *dv;
        while (PyDict_Next(conns, &pos, &dk, &dv)) {
            long a, b;
            if (!PyTuple_Check(dk) || PyTuple_GET_SIZE(dk) != 2 || long_of(PyTuple_GET_ITEM(dk, 0), &a) || long_of(PyTuple_GET_ITEM(dk, 1), &b)) {
                PyErr_Clear(); c.status = ST_PYTHON; free(pa); goto done;
            }
            pa[i] = key_id(&c, a, PyTuple_GET_ITEM(dk, 0));
            pb[i] = key_id(&c, b, PyTuple_GET_ITEM(dk, 1));
            if (pa[i] < 0 || pb[i] < 0) { PyErr_NoMemory(); c.status = ST_ERROR; free(pa); goto done; }
            ++i;
        }
    }
    /* required_for_output: layer by layer from the outputs backwards; stops at a layer that holds only input pins */
    frontier = (int*)malloc(sizeof(int) * (size_t)(c.nk + 1));
    layer = (int*)malloc(sizeof(int) * (size_t)(c.nk + 1));
    if (!frontier || !layer) { PyErr_NoMemory(); c.status = ST_ERROR; free(pa); goto done; }
    {
        int nf = 0;
        for (int i = 0; i < c.nk; ++i) if (c.k[i].is_out) frontier[nf++] = i;
        char* in_front = (char*)calloc((size_t)c.nk + 1, 1);
        char* in_layer = (char*)calloc((size_t)c.nk + 1, 1);
        if (!in_front || !in_layer) { free(in_front); free(in_layer); PyErr_NoMemory(); c.status = ST_ERROR; free(pa); goto done; }
        for (;;) {
            for (int i = 0; i < nf; ++i) in_front[frontier[i]] = 1;
            int nl = 0, hidden = 0;
            for (Py_ssize_t i = 0; i < n_conn; ++i)
                if (in_front[pb[i]] && !c.k[pa[i]].seen && !in_layer[pa[i]]) { in_layer[pa[i]] = 1; layer[nl++] = pa[i]; if (!c.k[pa[i]].is_in) ++hidden; }
            for (int i = 0; i < nf; ++i) in_front[frontier[i]] = 0;
            if (!nl || !hidden) { for (int i = 0; i < nl; ++i) in_layer[layer[i]] = 0; break; }
            for (int i = 0; i < nl; ++i) {
                in_layer[layer[i]] = 0;
                c.k[layer[i]].seen = 1;
                if (!c.k[layer[i]].is_in) c.k[layer[i]].needed = 1;
                frontier[i] = layer[i];
            }
            nf = nl;
        }
        free(in_front); free(in_layer);
    }
    free(pa);
    /* pass 2: incoming lists from the connection genes (cg.key, cg.weight, cg.enabled), insertion order */
    {
        Py_ssize_t pos = 0;
        PyObject *dk, *dv;
        while (PyDict_Next(conns, &pos, &dk, &dv)) {
            PyObject* en = PyObject_GetAttr(dv, s_enabled);
            if (!en) { c.status = ST_ERROR; goto done; }
            const int enabled = PyObject_IsTrue(en);
            Py_DECREF(en);
            if (enabled < 0) { c.status = ST_ERROR; goto done; }
            if (!enabled) continue;
            PyObject* key = PyObject_GetAttr(dv, s_key);
            if (!key) { c.status = ST_ERROR; goto done; }
            long a, b;
            if (!PyTuple_Check(key) || PyTuple_GET_SIZE(key) != 2 || long_of(PyTuple_GET_ITEM(key, 0), &a) || long_of(PyTuple_GET_ITEM(key, 1), &b)) {
                Py_DECREF(key); PyErr_Clear(); c.status = ST_PYTHON; goto done;
            }
            /* the key objects stay alive through the gene / the dict key (equal ints hash to one KeyInfo anyway) */
            const int src = key_id(&c, a, PyTuple_GET_ITEM(dk, 0)), dst = key_id(&c, b, PyTuple_GET_ITEM(dk, 1));
            Py_DECREF(key);
            if (src < 0 || dst < 0) { PyErr_NoMemory(); c.status = ST_ERROR; goto done; }
            if (!c.k[dst].needed && !c.k[src].needed) continue;
            if (c.k[src].is_out) continue;
            double w;
            if (attr_double(dv, s_weight, &w)) { c.status = ST_ERROR; goto done; }
            if (grow((void**)&c.e, &c.cape, c.ne + 1, sizeof(Edge))) { PyErr_NoMemory(); c.status = ST_ERROR; goto done; }
            c.e[c.ne].src = src; c.e[c.ne].w = w; c.e[c.ne].next = -1;
            if (c.k[dst].in_tail >= 0) c.e[c.k[dst].in_tail].next = c.ne; else c.k[dst].in_head = c.ne;
            c.k[dst].in_tail = c.ne++;
            c.k[dst].has_incoming = 1;
            c.k[src].has_incoming = 1;   /* incoming.setdefault(src, []) */
        }
    }
    /* the connection-dict key of an equal int may be a different object than the nodes-dict key: look nodes up by value */
    int* outs = (int*)malloc(sizeof(int) * (size_t)(n_used + 2));
    if (!outs) { PyErr_NoMemory(); c.status = ST_ERROR; goto done; }
    for (Py_ssize_t i = 0; i < n_used; ++i) {
        long ok = 0;
        long_of(PyList_GET_ITEM(out_keys, i), &ok);
        const int id = key_id(&c, ok, PyList_GET_ITEM(out_keys, i));
        if (visit(&c, id, 0)) { free(outs); goto done; }
        if (c.k[id].memo == MEMO_SLOT) outs[i] = c.k[id].slot;
        else {   /* constant output plane: identity node 1.0*(c*1.0)+0.0; bit 30: the plane is float32 in the reference */
            TermRec t; t.w = (double)c.k[id].cval; t.slot = SLOT_ONE;
            const int slot = emit(&c, ACT_IDENTITY, 0, &t, 1, 0.0, 1.0);
            if (slot < 0) { free(outs); goto done; }
            outs[i] = slot | OUT_F32_CONST;
        }
    }
    {
        const Py_ssize_t n_outs_padded = n_used + (n_used & 1);
        const Py_ssize_t bytes = 16 + 4 * n_outs_padded + 32 * (Py_ssize_t)c.nn + 16 * (Py_ssize_t)c.nt;
        PyObject* blob = PyBytes_FromStringAndSize(NULL, bytes);
        if (!blob) { free(outs); c.status = ST_ERROR; goto done; }
        unsigned char* p = (unsigned char*)PyBytes_AS_STRING(blob);
        int32_t hdr[4] = {BLOB_MAGIC, c.nn, c.nt, (int32_t)n_used};
        memcpy(p, hdr, 16); p += 16;
        for (Py_ssize_t i = 0; i < n_outs_padded; ++i) { int32_t v = i < n_used ? outs[i] : 0; memcpy(p, &v, 4); p += 4; }
        for (int i = 0; i < c.nn; ++i) {
            int32_t h4[4] = {c.nodes[i].act, c.nodes[i].agg, c.nodes[i].t0, c.nodes[i].nt};
            memcpy(p, h4, 16); memcpy(p + 16, &c.nodes[i].bias, 8); memcpy(p + 24, &c.nodes[i].resp, 8); p += 32;
        }
        for (int i = 0; i < c.nt; ++i) { int32_t s2[2] = {c.terms[i].slot, 0}; memcpy(p, &c.terms[i].w, 8); memcpy(p + 8, s2, 8); p += 16; }
        result = Py_BuildValue("(Ni)", blob, SLOT_NODE0 + c.nn);
    }
    free(outs);
done:
    free(frontier); free(layer);
    Py_XDECREF(conns); Py_XDECREF(nodes);
    const int status = c.status;
    ctx_free(&c);
    if (result) return result;
    if (status == ST_PYTHON && !PyErr_Occurred()) Py_RETURN_NONE;
    if (!PyErr_Occurred()) PyErr_SetString(PyExc_RuntimeError, "_flatten.flatten failed");
    return NULL;
}

static PyMethodDef methods[] = {
    {"flatten", py_flatten, METH_VARARGS, "flatten(genome, input_keys, output_keys, n_outputs) -> (bytes, n_slots) | None"},
    {NULL, NULL, 0, NULL}};
static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_flatten", "genome flattener (see genome.py)", -1, methods};
PyMODINIT_FUNC PyInit__flatten(void) {
    s_bias = PyUnicode_InternFromString("bias"); s_response = PyUnicode_InternFromString("response");
    s_aggregation = PyUnicode_InternFromString("aggregation"); s_activation = PyUnicode_InternFromString("activation");
    s_enabled = PyUnicode_InternFromString("enabled"); s_key = PyUnicode_InternFromString("key");
    s_weight = PyUnicode_InternFromString("weight"); s_connections = PyUnicode_InternFromString("connections");
    s_nodes = PyUnicode_InternFromString("nodes");
    if (!s_bias || !s_response || !s_aggregation || !s_activation || !s_enabled || !s_key || !s_weight || !s_connections || !s_nodes) return NULL;
    return PyModule_Create(&moduledef);
}
