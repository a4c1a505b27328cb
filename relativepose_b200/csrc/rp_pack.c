/* rp_pack.c -- host-side packing of a list of primitive-cache records (the reference's list of dicts,
 * trainRelativePoseModuleRecFD.py:207-208) into the concatenated arrays rp_solve_batch takes.
 *
 * RelativePoseEstimation_batch(list_of_dicts) has to gather 8 arrays per record -- 32 768 small arrays for a 4 096-pair batch.
 * numpy.concatenate spends ~1.5 us of bookkeeping per input array on top of the copy; here one C loop per field takes the
 * arrays through the buffer protocol (GIL held, ~0.2 us each) and then copies them with the GIL RELEASED, so the eight fields
 * pack concurrently on the caller's thread pool.  Pure host code: CPython C API + memcpy, no CUDA, no numpy C API.
 *
 *   pack_field(records, key, dst, fmt, itemsize, cols) -> bytes (int64 row count of every record)  |  None
 *
 * records: list of dicts; key: str; dst: writable C-contiguous buffer with room for all rows; fmt: 'd' or 'f' (the element
 * type every record must already have -- no conversion here); cols: columns per row (0: 1-D arrays).  Arrays that are not
 * C-contiguous (the reference's pipeline passes transposed descriptor views, rpmodule.py:531-532) are copied element-wise in
 * C order.  None = some record does not fit this fast path (other dtype, odd shape, not a buffer): the caller falls back to
 * numpy.concatenate, which converts. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

static int fmt_matches(const char* f, char want) {
    if (!f) return want == 'B';
    while (*f == '<' || *f == '=' || *f == '@' || *f == '|') ++f;          /* native / little-endian prefixes */
    return f[0] == want && f[1] == '\0';
}

static PyObject* pack_field(PyObject* self, PyObject* args) {
    PyObject *records, *key, *dst_obj;
    const char* fmt;
    Py_ssize_t itemsize, cols;
    (void)self;
    if (!PyArg_ParseTuple(args, "O!UOsnn", &PyList_Type, &records, &key, &dst_obj, &fmt, &itemsize, &cols)) return NULL;
    if (itemsize < 1 || cols < 0 || strlen(fmt) != 1) { PyErr_SetString(PyExc_ValueError, "bad itemsize / cols / fmt"); return NULL; }
    const Py_ssize_t B = PyList_GET_SIZE(records);
    Py_buffer dst;
    if (PyObject_GetBuffer(dst_obj, &dst, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) != 0) return NULL;
    Py_buffer* views = (Py_buffer*)PyMem_Calloc((size_t)(B > 0 ? B : 1), sizeof(Py_buffer));
    PyObject* counts = PyBytes_FromStringAndSize(NULL, B * (Py_ssize_t)sizeof(int64_t));
    if (!views || !counts) { PyBuffer_Release(&dst); PyMem_Free(views); Py_XDECREF(counts); return PyErr_NoMemory(); }
    int64_t* cnt = (int64_t*)PyBytes_AS_STRING(counts);
    const Py_ssize_t rowbytes = itemsize * (cols > 0 ? cols : 1);
    Py_ssize_t got = 0, total = 0;
    int ok = 1;
    for (Py_ssize_t i = 0; i < B && ok; ++i) {                               /* phase 1 (GIL held): take the buffers */
        PyObject* rec = PyList_GET_ITEM(records, i);
        PyObject* item = PyDict_Check(rec) ? PyDict_GetItemWithError(rec, key) : NULL;      /* borrowed */
        if (!item) { if (PyErr_Occurred()) PyErr_Clear(); ok = 0; break; }
        if (PyObject_GetBuffer(item, &views[i], PyBUF_RECORDS_RO) != 0) { PyErr_Clear(); ok = 0; break; }
        ++got;
        const Py_buffer* v = &views[i];
        if (v->itemsize != itemsize || !fmt_matches(v->format, fmt[0]) || v->len % rowbytes != 0) { ok = 0; break; }
        if (cols > 0 ? !((v->ndim == 2 && v->shape[1] == cols) || (v->ndim == 1)) : (v->ndim > 2)) { ok = 0; break; }
        cnt[i] = (int64_t)(v->len / rowbytes);
        total += v->len;
    }
    if (ok && total > dst.len) ok = 0;
    if (ok) {
        /* non-contiguous inputs need the interpreter-side helper: copy them now, mark them done */
        Py_ssize_t off = 0;
        for (Py_ssize_t i = 0; i < B && ok; ++i) {
            if (!PyBuffer_IsContiguous(&views[i], 'C')) {
                if (PyBuffer_ToContiguous((char*)dst.buf + off, &views[i], views[i].len, 'C') != 0) { PyErr_Clear(); ok = 0; }
                off += views[i].len;
                views[i].len = -views[i].len - 1;                             /* negative: already copied (len recoverable) */
            } else off += views[i].len;
        }
    }
    if (ok) {
        char* out = (char*)dst.buf;
        Py_BEGIN_ALLOW_THREADS                                               /* phase 2 (GIL released): the copies */
        for (Py_ssize_t i = 0; i < B; ++i) {
            Py_ssize_t len = views[i].len;
            if (len < 0) { out += -(len + 1); continue; }
            memcpy(out, views[i].buf, (size_t)len);
            out += len;
        }
        Py_END_ALLOW_THREADS
    }
    for (Py_ssize_t i = 0; i < got; ++i) {
        if (views[i].len < 0) views[i].len = -(views[i].len + 1);
        PyBuffer_Release(&views[i]);
    }
    PyMem_Free(views);
    PyBuffer_Release(&dst);
    if (!ok) { Py_DECREF(counts); Py_RETURN_NONE; }
    return counts;
}

static PyMethodDef methods[] = {
    {"pack_field", pack_field, METH_VARARGS, "pack_field(records, key, dst, fmt, itemsize, cols) -> bytes of int64 row counts, or None"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_rp_pack", "host-side record packing for relativepose_b200", -1, methods,
                                       NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__rp_pack(void) { return PyModule_Create(&moduledef); }
