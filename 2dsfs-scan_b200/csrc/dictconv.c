/* dictconv.c -- host-side adapter: the reference's data_dict (scripts/src/twoDSFS_class.py:132-134) -> flat arrays, in C.
 *
 * Called through ctypes.PyDLL (GIL held, Python exceptions propagate).  Semantics follow what every reference scanner does
 * per SNP: key "CHR-POS" split on '-' into exactly two parts (:176), pos = int(POS), calls.get(pop, (0, 0)) (:190-191),
 * annotation = snp_info.get('annotation') (:185).  Only the per-SNP Python loop is replaced; sorting happens in numpy.
 * Build: gcc -O2 -shared -fPIC -I<python include> dictconv.c   (see build.py) */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

static int id_of(PyObject* map, PyObject* names, PyObject* key) {
  PyObject* v = PyDict_GetItemWithError(map, key);
  if (v) return (int)PyLong_AsLong(v);
  if (PyErr_Occurred()) return -1;
  Py_ssize_t n = PyList_GET_SIZE(names);
  PyObject* idx = PyLong_FromSsize_t(n);
  if (!idx || PyDict_SetItem(map, key, idx) < 0 || PyList_Append(names, key) < 0) { Py_XDECREF(idx); return -1; }
  Py_DECREF(idx);
  return (int)n;
}

static int pair_of(PyObject* calls, PyObject* pop, int64_t* ref, int64_t* alt) {
  PyObject* c = PyDict_Check(calls) ? PyDict_GetItemWithError(calls, pop) : NULL;
  if (!c) {
    if (PyErr_Occurred()) return -1;
    if (!PyDict_Check(calls)) {  /* generic mapping: calls.get(pop, (0, 0)) */
      c = PyObject_CallMethod(calls, "get", "O(ii)", pop, 0, 0);
      if (!c) return -1;
      PyObject* a = PySequence_GetItem(c, 0);
      PyObject* b = a ? PySequence_GetItem(c, 1) : NULL;
      Py_DECREF(c);
      if (!a || !b) { Py_XDECREF(a); Py_XDECREF(b); return -1; }
      *ref = PyLong_AsLongLong(a); *alt = PyLong_AsLongLong(b);
      Py_DECREF(a); Py_DECREF(b);
      return PyErr_Occurred() ? -1 : 0;
    }
    *ref = 0; *alt = 0;
    return 0;
  }
  PyObject* a = PySequence_GetItem(c, 0);
  PyObject* b = a ? PySequence_GetItem(c, 1) : NULL;
  if (!a || !b) { Py_XDECREF(a); Py_XDECREF(b); return -1; }
  *ref = PyLong_AsLongLong(a); *alt = PyLong_AsLongLong(b);
  Py_DECREF(a); Py_DECREF(b);
  return PyErr_Occurred() ? -1 : 0;
}

/* pos[n], cnt[n*4], chrom_id[n], ann_id[n] are caller-allocated; chrom_names / ann_vocab are empty lists filled here. */
int tdsfs_dict_to_arrays(PyObject* d, PyObject* pop1, PyObject* pop2, int64_t* pos, int64_t* cnt, int32_t* chrom_id,
                         int32_t* ann_id, PyObject* chrom_names, PyObject* ann_vocab) {
  if (!PyDict_Check(d)) { PyErr_SetString(PyExc_TypeError, "data_dict must be a dict"); return -1; }
  PyObject* cmap = PyDict_New();
  PyObject* amap = PyDict_New();
  PyObject *s_calls = PyUnicode_InternFromString("calls"), *s_ann = PyUnicode_InternFromString("annotation");
  if (!cmap || !amap || !s_calls || !s_ann) goto fail;
  {
    Py_ssize_t it = 0, i = 0;
    PyObject *key, *val;
    const char* last_chrom = NULL; Py_ssize_t last_len = -1; int last_id = -1;
    while (PyDict_Next(d, &it, &key, &val)) {
      Py_ssize_t klen;
      const char* k = PyUnicode_AsUTF8AndSize(key, &klen);
      if (!k) goto fail;
      const char* dash = memchr(k, '-', (size_t)klen);
      if (!dash) { PyErr_SetString(PyExc_ValueError, "not enough values to unpack (expected 2, got 1)"); goto fail; }
      if (memchr(dash + 1, '-', (size_t)(klen - (dash + 1 - k)))) { PyErr_SetString(PyExc_ValueError, "too many values to unpack (expected 2)"); goto fail; }
      const Py_ssize_t clen = dash - k;
      if (!(last_chrom && clen == last_len && memcmp(k, last_chrom, (size_t)clen) == 0)) {
        PyObject* cs = PyUnicode_FromStringAndSize(k, clen);
        if (!cs) goto fail;
        last_id = id_of(cmap, chrom_names, cs);
        Py_DECREF(cs);
        if (last_id < 0) goto fail;
        last_chrom = k; last_len = clen;  /* k stays valid while `key` is alive (the dict holds it) */
      }
      chrom_id[i] = last_id;
      /* pos = int(POS): digits fast path, Python's own parser otherwise */
      const char* p = dash + 1; const Py_ssize_t plen = klen - clen - 1;
      int64_t v = 0; int fast = plen > 0 && plen < 18;
      for (Py_ssize_t q = 0; fast && q < plen; ++q) { if (p[q] < '0' || p[q] > '9') fast = 0; else v = v * 10 + (p[q] - '0'); }
      if (!fast) {
        PyObject* ps = PyUnicode_FromStringAndSize(p, plen);
        PyObject* pl = ps ? PyLong_FromUnicodeObject(ps, 10) : NULL;
        Py_XDECREF(ps);
        if (!pl) goto fail;
        v = PyLong_AsLongLong(pl);
        Py_DECREF(pl);
        if (PyErr_Occurred()) goto fail;
      }
      pos[i] = v;
      PyObject* calls = PyObject_GetItem(val, s_calls);  /* snp_info['calls'] : KeyError when absent */
      if (!calls) goto fail;
      int rc = pair_of(calls, pop1, &cnt[4 * i], &cnt[4 * i + 1]);
      if (!rc) rc = pair_of(calls, pop2, &cnt[4 * i + 2], &cnt[4 * i + 3]);
      Py_DECREF(calls);
      if (rc) goto fail;
      PyObject* ann = PyDict_Check(val) ? PyDict_GetItemWithError(val, s_ann) : NULL;  /* .get('annotation') -> None if absent */
      if (!ann) { if (PyErr_Occurred()) goto fail; ann = Py_None; }
      int aid = id_of(amap, ann_vocab, ann);
      if (aid < 0) goto fail;
      ann_id[i] = aid;
      ++i;
    }
  }
  Py_DECREF(cmap); Py_DECREF(amap); Py_DECREF(s_calls); Py_DECREF(s_ann);
  return 0;
fail:
  Py_XDECREF(cmap); Py_XDECREF(amap); Py_XDECREF(s_calls); Py_XDECREF(s_ann);
  if (!PyErr_Occurred()) PyErr_SetString(PyExc_RuntimeError, "tdsfs_dict_to_arrays failed");
  return -1;
}
