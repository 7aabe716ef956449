/* lib_wfst_pytargets.so — walks the reference's target argument, a Python list of lists of
 * ints (criterions/ctc.py:32 `targets`, benchmarks/ctc_benchmark.py:23-24), straight into the
 * int32 staging buffer the C ABI takes (include/wfst_b200.h: targets + target_offsets).
 *
 * In Python this is np.fromiter(itertools.chain.from_iterable(targets)): one interpreter
 * round trip per label, 1.6 ms for the 45 k labels of a B=256, L=176 batch — several times
 * the kernel it feeds.  Here it is a pointer walk (~0.1 ms).  Loaded with ctypes.PyDLL (the
 * GIL is held, arguments are borrowed references).  Kept apart from libwfst_b200.so so that
 * the C ABI library has no Python dependency.
 *
 *   wfst_pytargets_lengths(targets, lengths[B])        -> total number of labels, or -1 if
 *       `targets` is not a list/tuple of lists/tuples (the caller then takes its generic path)
 *   wfst_pytargets_fill(targets, out[total], minmax[2]) -> total, or -1 (element not an int /
 *       does not fit int32 / lengths changed); minmax receives the smallest and largest label
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

#define EXPORT __attribute__((visibility("default")))

static int is_seq(PyObject* o) { return PyList_CheckExact(o) || PyTuple_CheckExact(o); }

EXPORT long long wfst_pytargets_lengths(PyObject* targets, int32_t* lengths, long long B) {
  if (!targets || !is_seq(targets)) return -1;
  const Py_ssize_t n = PySequence_Fast_GET_SIZE(targets);
  if ((long long)n != B) return -1;
  PyObject** rows = PySequence_Fast_ITEMS(targets);
  long long total = 0;
  for (Py_ssize_t b = 0; b < n; ++b) {
    if (!is_seq(rows[b])) return -1;
    const Py_ssize_t len = PySequence_Fast_GET_SIZE(rows[b]);
    if (len > INT32_MAX) return -1;
    lengths[b] = (int32_t)len;
    total += len;
  }
  return total;
}

EXPORT long long wfst_pytargets_fill(PyObject* targets, int32_t* out, long long capacity, int32_t* minmax) {
  if (!targets || !is_seq(targets)) return -1;
  const Py_ssize_t n = PySequence_Fast_GET_SIZE(targets);
  PyObject** rows = PySequence_Fast_ITEMS(targets);
  long long k = 0;
  long lo = INT32_MAX, hi = INT32_MIN;
  for (Py_ssize_t b = 0; b < n; ++b) {
    if (!is_seq(rows[b])) return -1;
    const Py_ssize_t len = PySequence_Fast_GET_SIZE(rows[b]);
    if (k + len > capacity) return -1;
    PyObject** it = PySequence_Fast_ITEMS(rows[b]);
    for (Py_ssize_t i = 0; i < len; ++i) {
      PyObject* o = it[i];
      long v;
      if (!PyLong_CheckExact(o)) return -1;           /* bools, numpy scalars, tensors: generic path */
#if PY_VERSION_HEX >= 0x030C0000
      if (PyUnstable_Long_IsCompact((PyLongObject*)o)) {
        v = (long)PyUnstable_Long_CompactValue((PyLongObject*)o);
      } else
#endif
      {
        int overflow = 0;
        v = PyLong_AsLongAndOverflow(o, &overflow);
        if (overflow || (v == -1 && PyErr_Occurred())) { PyErr_Clear(); return -1; }
      }
      if (v < INT32_MIN || v > INT32_MAX) return -1;
      if (v < lo) lo = v;
      if (v > hi) hi = v;
      out[k++] = (int32_t)v;
    }
  }
  minmax[0] = (int32_t)(k ? lo : 0);
  minmax[1] = (int32_t)(k ? hi : 0);
  return k;
}
