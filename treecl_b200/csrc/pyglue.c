/* pyglue.c -- optional CPython helper for the ctypes binding (treecl_b200/_lib.py).
 *
 * td_encode_newick takes a table of C strings.  Built from Python with ctypes alone, that table costs a join + encode (two copies
 * of all Newick text) or one bytes object per tree; this helper hands the interpreter's own UTF-8 buffers of the str objects to
 * the library instead (PyUnicode_AsUTF8AndSize: no copy for ASCII strings) and releases the GIL around the call.
 * It does not link against libtreedist_cuda: the binding passes the address of td_encode_newick.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

typedef int (*encode_fn)(const char *const *, int64_t, int, void **);

static PyObject *encode_list(PyObject *self, PyObject *args)
{
    PyObject *list;
    unsigned long long fn_addr;
    int n_threads;
    (void)self;
    if (!PyArg_ParseTuple(args, "OKi", &list, &fn_addr, &n_threads)) return NULL;
    PyObject *seq = PySequence_Fast(list, "expected a sequence of Newick strings");
    if (!seq) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    const char **ptrs = (const char **)PyMem_Malloc(sizeof(char *) * (size_t)(n ? n : 1));
    if (!ptrs) { Py_DECREF(seq); return PyErr_NoMemory(); }
    Py_ssize_t held = 0;
    for (Py_ssize_t i = 0; i < n; i++, held++) {
        PyObject *item = PySequence_Fast_GET_ITEM(seq, i);
        Py_ssize_t len = 0;
        const char *p = NULL;
        if (PyUnicode_Check(item)) p = PyUnicode_AsUTF8AndSize(item, &len);
        else if (PyBytes_Check(item)) { p = PyBytes_AS_STRING(item); len = PyBytes_GET_SIZE(item); }
        else PyErr_SetString(PyExc_TypeError, "Newick strings must be str or bytes");
        if (p && strlen(p) != (size_t)len) { PyErr_SetString(PyExc_ValueError, "Newick strings must not contain NUL characters"); p = NULL; }
        if (!p) break;
        Py_INCREF(item);                     /* the buffers stay valid while the GIL is released */
        ptrs[i] = p;
    }
    PyObject *result = NULL;
    if (held == n) {
        void *ts = NULL;
        int rc;
        Py_BEGIN_ALLOW_THREADS
        rc = ((encode_fn)(uintptr_t)fn_addr)(ptrs, (int64_t)n, n_threads, &ts);
        Py_END_ALLOW_THREADS
        result = Py_BuildValue("iK", rc, (unsigned long long)(uintptr_t)ts);
    }
    for (Py_ssize_t i = 0; i < held; i++) Py_DECREF(PySequence_Fast_GET_ITEM(seq, i));
    PyMem_Free((void *)ptrs);
    Py_DECREF(seq);
    return result;
}

static PyMethodDef methods[] = {
    {"encode_list", encode_list, METH_VARARGS, "encode_list(strings, address_of_td_encode_newick, n_threads) -> (status, handle)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_pyglue", "zero-copy marshalling of Newick strings for libtreedist_cuda", -1, methods,
                                    NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__pyglue(void) { return PyModule_Create(&module); }
