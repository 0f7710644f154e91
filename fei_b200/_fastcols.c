/* Host-side marshalling helper for MemoryChain.validate_chain (fei_b200/memdir_tools/memorychain.py):
 * reads the hashed fields of a list of block objects straight out of their instance dicts and returns them as
 * typed columns (UTF-8 blob + offsets / int64 / float64 / all-None), the shape libfeiscan's fei_json_col wants.
 * Pure CPython C API, no numpy.  Anything unusual (missing key, mixed types, subclasses, big ints, lone surrogates)
 * makes it return None and the pure-Python path in memorychain.py takes over: this is glue, not a compute path. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

enum { K_STR = 1, K_INT = 2, K_FLOAT = 3, K_NONE = 4 };

/* column from n borrowed values; returns new ref tuple (kind, payload...) or NULL with *soft = 1 for "fall back" */
static PyObject* build_column(PyObject** v, Py_ssize_t n, int* soft) {
  int kind = 0;
  for (Py_ssize_t i = 0; i < n; ++i) {
    PyObject* o = v[i];
    int k = PyUnicode_CheckExact(o) ? K_STR : PyLong_CheckExact(o) ? K_INT : PyFloat_CheckExact(o) ? K_FLOAT : o == Py_None ? K_NONE : -1;
    if (k < 0 || (kind && k != kind)) { *soft = 1; return NULL; }
    kind = k;
  }
  if (n == 0) kind = K_NONE;
  if (kind == K_NONE) return Py_BuildValue("(i)", K_NONE);
  if (kind == K_INT) {
    PyObject* b = PyBytes_FromStringAndSize(NULL, n * 8);
    if (!b) return NULL;
    int64_t* out = (int64_t*)PyBytes_AS_STRING(b);
    for (Py_ssize_t i = 0; i < n; ++i) {
      int ovf = 0;
      long long x = PyLong_AsLongLongAndOverflow(v[i], &ovf);
      if (ovf) { Py_DECREF(b); *soft = 1; return NULL; }
      out[i] = (int64_t)x;
    }
    return Py_BuildValue("(iN)", K_INT, b);
  }
  if (kind == K_FLOAT) {
    PyObject* b = PyBytes_FromStringAndSize(NULL, n * 8);
    if (!b) return NULL;
    double* out = (double*)PyBytes_AS_STRING(b);
    for (Py_ssize_t i = 0; i < n; ++i) out[i] = PyFloat_AS_DOUBLE(v[i]);
    return Py_BuildValue("(iN)", K_FLOAT, b);
  }
  /* strings: offsets first (total size), then the blob */
  PyObject* offs = PyBytes_FromStringAndSize(NULL, (n + 1) * 8);
  if (!offs) return NULL;
  uint64_t* off = (uint64_t*)PyBytes_AS_STRING(offs);
  off[0] = 0;
  for (Py_ssize_t i = 0; i < n; ++i) {
    Py_ssize_t len;
    const char* s = PyUnicode_AsUTF8AndSize(v[i], &len);       /* cached on the str object; fails on lone surrogates */
    if (!s) { PyErr_Clear(); Py_DECREF(offs); *soft = 1; return NULL; }
    off[i + 1] = off[i] + (uint64_t)len;
  }
  PyObject* blob = PyBytes_FromStringAndSize(NULL, (Py_ssize_t)off[n]);
  if (!blob) { Py_DECREF(offs); return NULL; }
  char* dst = PyBytes_AS_STRING(blob);
  for (Py_ssize_t i = 0; i < n; ++i) {
    Py_ssize_t len;
    const char* s = PyUnicode_AsUTF8AndSize(v[i], &len);
    memcpy(dst + off[i], s, (size_t)len);
  }
  return Py_BuildValue("(iNN)", K_STR, blob, offs);
}

/* columns(blocks: list, names: tuple[str, ...], memory_id_at: int) -> tuple | None
 * names are instance-dict keys; the entry at index memory_id_at is "memory_data" and is replaced by
 * memory_data.get("metadata", {}).get("unique_id", "")  (memorychain.py:120). */
static PyObject* columns(PyObject* self, PyObject* args) {
  PyObject *blocks, *names;
  Py_ssize_t mid_at;
  if (!PyArg_ParseTuple(args, "O!O!n", &PyList_Type, &blocks, &PyTuple_Type, &names, &mid_at)) return NULL;
  const Py_ssize_t n = PyList_GET_SIZE(blocks), nf = PyTuple_GET_SIZE(names);
  PyObject* k_meta = PyUnicode_InternFromString("metadata");
  PyObject* k_uid = PyUnicode_InternFromString("unique_id");
  PyObject* empty = PyUnicode_InternFromString("");
  PyObject** vals = (PyObject**)PyMem_Malloc(sizeof(PyObject*) * (size_t)(n ? n : 1) * (size_t)nf);
  PyObject* result = NULL;
  int soft = 0;
  if (!vals || !k_meta || !k_uid || !empty) { PyErr_NoMemory(); goto done; }
  for (Py_ssize_t i = 0; i < n && !soft; ++i) {
    PyObject* d = PyObject_GenericGetDict(PyList_GET_ITEM(blocks, i), NULL);      /* new reference */
    if (!d) { PyErr_Clear(); soft = 1; break; }
    if (!PyDict_CheckExact(d)) { Py_DECREF(d); soft = 1; break; }
    for (Py_ssize_t f = 0; f < nf; ++f) {
      PyObject* x = PyDict_GetItemWithError(d, PyTuple_GET_ITEM(names, f));       /* borrowed; the block keeps it alive */
      if (!x) { PyErr_Clear(); soft = 1; break; }
      if (f == mid_at) {
        if (!PyDict_CheckExact(x)) { soft = 1; break; }
        PyObject* meta = PyDict_GetItemWithError(x, k_meta);
        if (!meta) { if (PyErr_Occurred()) { PyErr_Clear(); soft = 1; break; } x = empty; }
        else if (!PyDict_CheckExact(meta)) { soft = 1; break; }
        else {
          x = PyDict_GetItemWithError(meta, k_uid);
          if (!x) { if (PyErr_Occurred()) { PyErr_Clear(); soft = 1; break; } x = empty; }
        }
      }
      vals[f * n + i] = x;
    }
    Py_DECREF(d);
  }
  if (!soft) {
    result = PyTuple_New(nf);
    for (Py_ssize_t f = 0; result && f < nf; ++f) {
      PyObject* col = build_column(vals + f * n, n, &soft);
      if (!col) { Py_CLEAR(result); break; }
      PyTuple_SET_ITEM(result, f, col);
    }
  }
  if (!result && soft && !PyErr_Occurred()) { result = Py_None; Py_INCREF(result); }
done:
  PyMem_Free(vals);
  Py_XDECREF(k_meta); Py_XDECREF(k_uid); Py_XDECREF(empty);
  return result;
}

static PyMethodDef methods[] = {{"columns", columns, METH_VARARGS, "typed columns of block instance attributes, or None"}, {NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fastcols", NULL, -1, methods};
PyMODINIT_FUNC PyInit__fastcols(void) { return PyModule_Create(&moddef); }
