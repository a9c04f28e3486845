/* swbhost.c — native helpers of the Python host layer (indelpost_b200/_swbhost*.so, a CPython extension).
 *
 * The reference's host layer is compiled Cython (sswpy.pyx); the per-pair work it does in C -- copying a read into an
 * int8 array (sswpy.pyx:149-170), turning an s_align into the 7-field Alignment tuple with its "%d%s" CIGAR string
 * (sswpy.pyx:283-298) -- is done here in C as well, for whole batches:
 *
 *   gather(seqs, dest, cap, off, len, byte0)      copy a sequence of str / bytes straight into a (pinned) staging blob (GIL released,
 *                                                 several threads for large tables) and fill the offset / length tables of swb_batch
 *   alignment(res, arena, k, cls)                 swb_result record k -> cls(CIGAR, score1, score2, ref_begin1, ref_end1,
 *   alignments(res, arena, k0, k1, cls)           read_begin1, read_end1)   (one, or a list for a range)
 *
 * Addresses are plain integers (numpy's .ctypes.data).  No GPU code here; libswb200.so stays Python-free.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>
#include <pthread.h>

/* same layout as swb_result (include/swb200.h) */
typedef struct {
    uint16_t score1, score2;
    int32_t ref_begin1, ref_end1, read_begin1, read_end1, ref_end2, cigar_len;
    uint16_t flag, status;
    int64_t cigar_off;
} swb_result_t;

/* gather: pass 1 (GIL held) collects (pointer, length) of every item and fills the offset / length tables; pass 2 copies the
 * bytes -- with the GIL released and, for large tables, split over a few threads by byte count (the objects are kept alive by the
 * caller's sequence, which `fast` references; bytes / compact-ASCII str buffers are immutable).
 * Returns the end byte, or -(needed capacity) - 1 when `cap` is too small (nothing copied: the caller grows its buffer and calls again). */
typedef struct { const char** ptr; const int64_t* off; const int32_t* len; char* dest; Py_ssize_t i0, i1; } gather_job;

static void* gather_worker(void* a)
{
    const gather_job* j = (const gather_job*)a;
    for (Py_ssize_t i = j->i0; i < j->i1; ++i) memcpy(j->dest + j->off[i], j->ptr[i], (size_t)j->len[i]);
    return NULL;
}

#define GATHER_MAX_THREADS 8

static PyObject* host_gather(PyObject* self, PyObject* args)
{
    PyObject* seqs;
    unsigned long long dest_a, off_a, len_a;
    long long cap, byte0;
    if (!PyArg_ParseTuple(args, "OKLKKL", &seqs, &dest_a, &cap, &off_a, &len_a, &byte0)) return NULL;
    PyObject* fast = PySequence_Fast(seqs, "gather: expected a sequence of str / bytes");
    if (!fast) return NULL;
    char* dest = (char*)(uintptr_t)dest_a;
    int64_t* off = (int64_t*)(uintptr_t)off_a;
    int32_t* len = (int32_t*)(uintptr_t)len_a;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject** items = PySequence_Fast_ITEMS(fast);
    const char** ptr = (const char**)PyMem_Malloc((size_t)(n > 0 ? n : 1) * sizeof(char*));
    if (!ptr) { Py_DECREF(fast); return PyErr_NoMemory(); }
    long long pos = byte0;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject* o = items[i];
        const char* p; Py_ssize_t l;
        if (PyBytes_Check(o)) { p = PyBytes_AS_STRING(o); l = PyBytes_GET_SIZE(o); }
        else if (PyUnicode_Check(o)) {
            if (PyUnicode_IS_COMPACT_ASCII(o)) { p = (const char*)PyUnicode_1BYTE_DATA(o); l = PyUnicode_GET_LENGTH(o); }
            else { p = PyUnicode_AsUTF8AndSize(o, &l); if (!p) { PyMem_Free(ptr); Py_DECREF(fast); return NULL; } }   /* obj_to_cstr_len encodes utf8, sswpy.pyx:45-55 (the utf8 copy is cached in the object) */
        } else { PyMem_Free(ptr); Py_DECREF(fast); PyErr_SetString(PyExc_TypeError, "expected str or bytes"); return NULL; }
        if (l > INT32_MAX) { PyMem_Free(ptr); Py_DECREF(fast); PyErr_SetString(PyExc_OverflowError, "sequence too long"); return NULL; }
        ptr[i] = p; off[i] = pos; len[i] = (int32_t)l;
        pos += l;
    }
    if (pos > cap) { PyMem_Free(ptr); Py_DECREF(fast); return PyLong_FromLongLong(-pos - 1); }
    const long long total = pos - byte0;
    int nt = total >= (8ll << 20) ? GATHER_MAX_THREADS : (total >= (1ll << 20) ? 2 : 1);
    if (nt > n) nt = n > 0 ? (int)n : 1;
    Py_BEGIN_ALLOW_THREADS
    gather_job jobs[GATHER_MAX_THREADS];
    pthread_t th[GATHER_MAX_THREADS];
    int started[GATHER_MAX_THREADS];
    Py_ssize_t i0 = 0;
    for (int t = 0; t < nt; ++t) {
        /* entries [i0, i1): up to the t+1-th share of the bytes */
        const long long until = byte0 + total * (t + 1) / nt;
        Py_ssize_t i1 = i0;
        if (t + 1 == nt) i1 = n;
        else while (i1 < n && off[i1] < until) ++i1;
        jobs[t].ptr = ptr; jobs[t].off = off; jobs[t].len = len; jobs[t].dest = dest; jobs[t].i0 = i0; jobs[t].i1 = i1;
        started[t] = 0;
        if (t + 1 < nt && i1 > i0) started[t] = pthread_create(&th[t], NULL, gather_worker, &jobs[t]) == 0;
        if (!started[t] && t + 1 < nt) gather_worker(&jobs[t]);
        i0 = i1;
    }
    gather_worker(&jobs[nt - 1]);                     /* the calling thread takes the last share */
    for (int t = 0; t + 1 < nt; ++t) if (started[t]) pthread_join(th[t], NULL);
    Py_END_ALLOW_THREADS
    PyMem_Free(ptr);
    Py_DECREF(fast);
    return PyLong_FromLongLong(pos);
}

/* total byte length of a sequence of str / bytes (sizes the staging blob) */
static PyObject* host_total_len(PyObject* self, PyObject* seqs)
{
    PyObject* fast = PySequence_Fast(seqs, "total_len: expected a sequence of str / bytes");
    if (!fast) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject** items = PySequence_Fast_ITEMS(fast);
    long long tot = 0;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject* o = items[i];
        if (PyBytes_Check(o)) tot += PyBytes_GET_SIZE(o);
        else if (PyUnicode_Check(o)) {
            if (PyUnicode_IS_COMPACT_ASCII(o)) tot += PyUnicode_GET_LENGTH(o);
            else { Py_ssize_t l; if (!PyUnicode_AsUTF8AndSize(o, &l)) { Py_DECREF(fast); return NULL; } tot += l; }
        } else { Py_DECREF(fast); PyErr_SetString(PyExc_TypeError, "expected str or bytes"); return NULL; }
    }
    Py_DECREF(fast);
    return PyLong_FromLongLong(tot);
}

/* "%d%s" per BAM-packed op, MAPSTR "MIDNSHP=X" (ssw.h:171-190; codes above 8 print as 'M') */
static PyObject* cigar_str(const uint32_t* ops, int32_t n)
{
    if (n <= 0) { Py_RETURN_NONE; }
    char stack[512];
    char* buf = stack;
    const size_t need = (size_t)n * 11 + 1;
    if (need > sizeof stack) { buf = (char*)PyMem_Malloc(need); if (!buf) return PyErr_NoMemory(); }
    char* w = buf;
    for (int32_t i = 0; i < n; ++i) {
        uint32_t v = ops[i] >> 4; const uint32_t op = ops[i] & 15u;
        char tmp[10]; int t = 0;
        do { tmp[t++] = (char)('0' + v % 10); v /= 10; } while (v);
        while (t) *w++ = tmp[--t];
        *w++ = op > 8 ? 'M' : "MIDNSHP=X"[op];
    }
    PyObject* s = PyUnicode_FromStringAndSize(buf, w - buf);
    if (buf != stack) PyMem_Free(buf);
    return s;
}

static PyObject* make_alignment(const swb_result_t* r, const uint32_t* arena, PyObject* cls)
{
    PyObject* cig = cigar_str(arena + r->cigar_off, r->cigar_len);
    if (!cig) return NULL;
    PyObject* t = Py_BuildValue("(Niiiiii)", cig, (int)r->score1, (int)r->score2, (int)r->ref_begin1, (int)r->ref_end1, (int)r->read_begin1, (int)r->read_end1);
    if (!t || cls == Py_None) return t;
    PyObject* mk = PyObject_GetAttrString(cls, "_make");           /* namedtuple constructor from an iterable */
    if (!mk) { Py_DECREF(t); return NULL; }
    PyObject* out = PyObject_CallOneArg(mk, t);
    Py_DECREF(mk); Py_DECREF(t);
    return out;
}

static PyObject* host_alignment(PyObject* self, PyObject* args)
{
    unsigned long long res_a, arena_a; long long k; PyObject* cls;
    if (!PyArg_ParseTuple(args, "KKLO", &res_a, &arena_a, &k, &cls)) return NULL;
    return make_alignment((const swb_result_t*)(uintptr_t)res_a + k, (const uint32_t*)(uintptr_t)arena_a, cls);
}

static PyObject* host_alignments(PyObject* self, PyObject* args)
{
    unsigned long long res_a, arena_a; long long k0, k1; PyObject* cls;
    if (!PyArg_ParseTuple(args, "KKLLO", &res_a, &arena_a, &k0, &k1, &cls)) return NULL;
    if (k1 < k0) k1 = k0;
    PyObject* mk = NULL;
    if (cls != Py_None) { mk = PyObject_GetAttrString(cls, "_make"); if (!mk) return NULL; }
    PyObject* out = PyList_New((Py_ssize_t)(k1 - k0));
    if (!out) { Py_XDECREF(mk); return NULL; }
    const swb_result_t* res = (const swb_result_t*)(uintptr_t)res_a;
    const uint32_t* arena = (const uint32_t*)(uintptr_t)arena_a;
    for (long long k = k0; k < k1; ++k) {
        const swb_result_t* r = res + k;
        PyObject* cig = cigar_str(arena + r->cigar_off, r->cigar_len);
        PyObject* t = cig ? Py_BuildValue("(Niiiiii)", cig, (int)r->score1, (int)r->score2, (int)r->ref_begin1, (int)r->ref_end1, (int)r->read_begin1, (int)r->read_end1) : NULL;
        if (t && mk) { PyObject* a = PyObject_CallOneArg(mk, t); Py_DECREF(t); t = a; }
        if (!t) { Py_DECREF(out); Py_XDECREF(mk); return NULL; }
        PyList_SET_ITEM(out, (Py_ssize_t)(k - k0), t);
    }
    Py_XDECREF(mk);
    return out;
}

static PyMethodDef methods[] = {
    {"gather", host_gather, METH_VARARGS, "gather(seqs, dest_addr, cap, off_addr, len_addr, byte0) -> end byte"},
    {"total_len", host_total_len, METH_O, "total_len(seqs) -> bytes"},
    {"alignment", host_alignment, METH_VARARGS, "alignment(res_addr, arena_addr, k, cls) -> cls instance"},
    {"alignments", host_alignments, METH_VARARGS, "alignments(res_addr, arena_addr, k0, k1, cls) -> list"},
    {NULL, NULL, 0, NULL},
};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_swbhost", "native helpers of indelpost_b200's host layer", -1, methods};

PyMODINIT_FUNC PyInit__swbhost(void) { return PyModule_Create(&moddef); }
