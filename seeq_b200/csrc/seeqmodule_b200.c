/*
 * seeqmodule_b200.c -- the reference's CPython module, compiled IN PLACE from the read-only mount
 * (no source copied: SQB_REFERENCE_MODULE is the path of /root/reference/src/seeqmodule.c, given on the
 * compiler command line by seeq_b200/build.py) plus ONE method added to its SeeqObject type:
 *
 *    SeeqObject.matchBatch(data, mode="best", raw=False)
 *
 * The reference's methods (seeqmodule.c:986-1014: match, matchBest, matchAll, matchIter, matchPrefix,
 * matchSuffix) make one seeqStringMatch call per string -- on a GPU that is one launch-latency-bound
 * round trip per read.  matchBatch hands a whole buffer of reads to seeqBatchMatch (seeq_api.c): one pass
 * of K1..K4 over all lines.
 *
 *    data   bytes / bytearray / memoryview of '\n'-separated reads, or a list / tuple of str or bytes
 *           (joined with '\n' here)
 *    mode   "first" | "best" | "all"       (SQ_FIRST / SQ_BEST / SQ_ALL, libseeq.h:34-37)
 *    raw    False: a list of (line, start, end, dist) tuples in file order, line 0-based; start / end /
 *           dist are match_t's fields (what SeeqMatch.matches holds, seeqmodule.c:330-352)
 *           True:  a bytes object of packed uint32 quadruples {line, start, end, dist}
 *                  (numpy.frombuffer(..., dtype=uint32).reshape(-1, 4))
 *
 * The object's non-DNA mode (compile(..., mode): SQ_CONVERT or SQ_IGNORE, seeqmodule.c:1071-1074) applies
 * as in every other method.  Errors raise seeq.clibexception like the reference's methods do.
 */
#define PyInit_seeq PyInit_seeq_reference_
#include SQB_REFERENCE_MODULE
#undef PyInit_seeq

#include "seeq_b200.h"

static PyObject *
SeeqObject_matchBatch(SeeqObject *self, PyObject *args, PyObject *kwds)
{
   static char *kwlist[] = {"data", "mode", "raw", NULL};
   PyObject *data = NULL;
   const char *mode = "best";
   int raw = 0;
   if (!PyArg_ParseTupleAndKeywords(args, kwds, "O|sp:matchBatch", kwlist, &data, &mode, &raw)) return NULL;
   if (self->sq == NULL) {
      PyErr_SetString(SeeqException, "NULL reference to DFA pointer");
      return NULL;
   }
   int match_opt;
   if (strcmp(mode, "first") == 0) match_opt = SQ_FIRST;
   else if (strcmp(mode, "best") == 0) match_opt = SQ_BEST;
   else if (strcmp(mode, "all") == 0) match_opt = SQ_ALL;
   else {
      PyErr_SetString(SeeqException, "mode must be 'first', 'best' or 'all'");
      return NULL;
   }

   /* the text: a buffer as it is, or the items of a sequence joined with '\n' */
   Py_buffer view;
   int have_view = 0;
   PyObject *joined = NULL;
   const char *text = NULL;
   Py_ssize_t nbytes = 0;
   if (PyObject_CheckBuffer(data)) {
      if (PyObject_GetBuffer(data, &view, PyBUF_SIMPLE) < 0) return NULL;
      have_view = 1;
      text = (const char *)view.buf;
      nbytes = view.len;
   } else if (PyList_Check(data) || PyTuple_Check(data)) {
      PyObject *fast = PySequence_Fast(data, "matchBatch: data must be a buffer or a sequence of strings");
      if (fast == NULL) return NULL;
      const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
      Py_ssize_t total = 0;
      for (Py_ssize_t i = 0; i < n; i++) {
         PyObject *it = PySequence_Fast_GET_ITEM(fast, i);
         Py_ssize_t len = 0;
         if (PyUnicode_Check(it)) {
            if (PyUnicode_AsUTF8AndSize(it, &len) == NULL) { Py_DECREF(fast); return NULL; }
         } else if (PyBytes_Check(it)) {
            len = PyBytes_GET_SIZE(it);
         } else {
            Py_DECREF(fast);
            PyErr_SetString(PyExc_TypeError, "matchBatch: items must be str or bytes");
            return NULL;
         }
         total += len + 1;
      }
      joined = PyBytes_FromStringAndSize(NULL, total);
      if (joined == NULL) { Py_DECREF(fast); return NULL; }
      char *dst = PyBytes_AS_STRING(joined);
      for (Py_ssize_t i = 0; i < n; i++) {
         PyObject *it = PySequence_Fast_GET_ITEM(fast, i);
         const char *src;
         Py_ssize_t len = 0;
         if (PyUnicode_Check(it)) src = PyUnicode_AsUTF8AndSize(it, &len);
         else { src = PyBytes_AS_STRING(it); len = PyBytes_GET_SIZE(it); }
         memcpy(dst, src, (size_t)len);
         dst[len] = '\n';
         dst += len + 1;
      }
      Py_DECREF(fast);
      text = PyBytes_AS_STRING(joined);
      nbytes = total;
   } else {
      PyErr_SetString(PyExc_TypeError, "matchBatch: data must be a buffer or a list / tuple of strings");
      return NULL;
   }

   const sqb_rec_t *recs = NULL;
   long n;
   Py_BEGIN_ALLOW_THREADS
   n = seeqBatchMatch(self->sq, text, (size_t)nbytes, match_opt | self->options, 0 /* SQ_ANY */, &recs, NULL);
   Py_END_ALLOW_THREADS
   if (have_view) PyBuffer_Release(&view);
   Py_XDECREF(joined);
   if (n < 0) {
      PyErr_SetString(LibSeeqException, seeqPrintError());
      return NULL;
   }
   if (raw) return PyBytes_FromStringAndSize((const char *)recs, (Py_ssize_t)n * (Py_ssize_t)sizeof(sqb_rec_t));
   PyObject *list = PyList_New((Py_ssize_t)n);
   if (list == NULL) return NULL;
   for (long i = 0; i < n; i++) {
      PyObject *t = Py_BuildValue("(kkkk)", (unsigned long)recs[i].line, (unsigned long)recs[i].start,
                                  (unsigned long)recs[i].end, (unsigned long)recs[i].dist);
      if (t == NULL) { Py_DECREF(list); return NULL; }
      PyList_SET_ITEM(list, (Py_ssize_t)i, t);
   }
   return list;
}

static PyMethodDef SeeqObject_matchBatch_def = {
   "matchBatch", (PyCFunction)(void (*)(void))SeeqObject_matchBatch, METH_VARARGS | METH_KEYWORDS,
   "matchBatch(data, mode='best', raw=False)\nMatches every line of a buffer of '\\n'-separated reads (or every "
   "string of a list) in ONE pass on the GPU.\nReturns: list of (line, start, end, dist) tuples in file order "
   "(line is 0-based), or, with raw=True, the same as packed uint32 quadruples in a bytes object."
};

PyMODINIT_FUNC
PyInit_seeq(void)
{
   PyObject *m = PyInit_seeq_reference_();
   if (m == NULL) return NULL;
   PyObject *descr = PyDescr_NewMethod(&SeeqObjectType, &SeeqObject_matchBatch_def);
   if (descr == NULL) { Py_DECREF(m); return NULL; }
#if PY_VERSION_HEX >= 0x030C0000
   PyObject *dict = PyType_GetDict(&SeeqObjectType);
#else
   PyObject *dict = SeeqObjectType.tp_dict;
   Py_XINCREF(dict);
#endif
   const int rc = dict ? PyDict_SetItemString(dict, "matchBatch", descr) : -1;
   Py_XDECREF(dict);
   Py_DECREF(descr);
   if (rc < 0) { Py_DECREF(m); return NULL; }
   PyType_Modified(&SeeqObjectType);
   return m;
}
