// pymarshal.cpp -- CPython extension `zksnake_b200._marshal`: list[int] <-> little-endian 64-bit limb arrays.
//
// The reference crosses its FFI with one BigUint conversion per element, serially, under the GIL
// (/root/reference/src/bn254/polynomial.rs:537-540 `coeffs.iter().map(|x| Fr::from(x.clone()))`, src/bn254/curve.rs:358-361 for
// MSM scalars, and back through `.into()` at polynomial.rs:544): at 2^20 elements that marshalling costs more than the GPU
// work it feeds.  Here the same boundary is crossed by reading the PyLong digit arrays directly, in parallel host threads:
// Python ints are immutable, the calling thread keeps the GIL (so the list cannot change and no object can die) while plain
// C++ threads read them.  Values that need arithmetic (negative ints, ints wider than the limb count: `Fr::from(BigUint)`
// reduces mod r) are rare and are finished one by one through the Python number protocol after the threads have joined.
//
//   ints_to_limbs(seq, addr, nlimbs, modulus=None, item=-1, allow_negative=True) -> int
//        count of elements written to the uint64 buffer at `addr` (len(seq) * nlimbs words; numpy array or pinned host
//        memory); item >= 0: the elements are tuples and the integer is element[item]
//   limbs_to_ints(addr, count, nlimbs) -> list[int]
//   set_threads(n) / get_threads()
//
// No CUDA in this file: it is host-side plumbing of the binding layer (g++ only), not a compute path.
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#if PY_VERSION_HEX < 0x03090000
#error "zksnake_b200._marshal needs CPython >= 3.9"
#endif

namespace {

int g_threads = 0;   // 0 = min(hardware_concurrency, 16)

int thread_count(size_t n) {
  int t = g_threads;
  if (t <= 0) {
    unsigned hw = std::thread::hardware_concurrency();
    t = hw ? (int)std::min(hw, 16u) : 4;
  }
  // below ~16K elements the thread start-up costs more than the conversion
  size_t cap = n / 16384 + 1;
  if ((size_t)t > cap) t = (int)cap;
  return t < 1 ? 1 : t;
}

// ---- PyLong internals (30-bit digits on every 64-bit CPython build) ---------------------------------------------------------
static_assert(PyLong_SHIFT == 30, "30-bit PyLong digits expected");

struct LongView {
  const digit* d;
  Py_ssize_t nd;   // number of digits (0 for the value 0)
  bool negative;
};

inline LongView view_long(PyObject* o) {
  const PyLongObject* v = (const PyLongObject*)o;
  LongView r;
#if PY_VERSION_HEX >= 0x030C0000
  const uintptr_t tag = v->long_value.lv_tag;
  r.d = v->long_value.ob_digit;
  r.nd = (Py_ssize_t)(tag >> 3);
  r.negative = (tag & 3) == 2;
  if ((tag & 3) == 1) r.nd = 0;   // zero
#else
  Py_ssize_t sz = Py_SIZE(v);
  r.d = v->ob_digit;
  r.nd = sz < 0 ? -sz : sz;
  r.negative = sz < 0;
#endif
  return r;
}

// returns false when the value does not fit nlimbs 64-bit words or is negative (-> slow path)
inline bool long_to_limbs(PyObject* o, uint64_t* out, int nlimbs) {
  LongView v = view_long(o);
  for (int i = 0; i < nlimbs; i++) out[i] = 0;
  if (v.nd == 0) return true;
  if (v.negative) return false;
  const Py_ssize_t bits = (v.nd - 1) * 30 + (32 - __builtin_clz((unsigned)v.d[v.nd - 1]));
  if (bits > (Py_ssize_t)nlimbs * 64) return false;
  for (Py_ssize_t i = 0; i < v.nd; i++) {
    const uint64_t dg = v.d[i];
    const Py_ssize_t pos = i * 30;
    const int limb = (int)(pos >> 6), off = (int)(pos & 63);
    out[limb] |= dg << off;
    if (off > 34 && limb + 1 < nlimbs) out[limb + 1] |= dg >> (64 - off);
  }
  return true;
}

// ---- ints_to_limbs -----------------------------------------------------------------------------------------------------------
PyObject* ints_to_limbs(PyObject*, PyObject* args) {
  PyObject *seq, *modulus = Py_None;
  unsigned long long addr;
  int nlimbs, item = -1, allow_negative = 1;
  if (!PyArg_ParseTuple(args, "OKi|Oip", &seq, &addr, &nlimbs, &modulus, &item, &allow_negative)) return nullptr;
  if (nlimbs < 1 || nlimbs > 16) {
    PyErr_SetString(PyExc_ValueError, "nlimbs must be in 1..16");
    return nullptr;
  }
  PyObject* fast = PySequence_Fast(seq, "ints_to_limbs expects a list or tuple of ints");
  if (!fast) return nullptr;
  const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
  PyObject** items = PySequence_Fast_ITEMS(fast);
  uint64_t* out = (uint64_t*)(uintptr_t)addr;
  if (n && !out) {
    Py_DECREF(fast);
    PyErr_SetString(PyExc_ValueError, "null output buffer");
    return nullptr;
  }
  const int nt = thread_count((size_t)n);
  std::vector<std::vector<Py_ssize_t>> slow(nt);
  std::atomic<Py_ssize_t> bad_type(-1);
  // Workers check the type and convert.  Every element is its own heap object, so the loop is a pointer chase: prefetching a
  // few objects ahead hides most of the cache misses, and the threads hide the rest.
  auto work = [&](int t) {
    const Py_ssize_t lo = n * t / nt, hi = n * (t + 1) / nt;
    for (Py_ssize_t i = lo; i < hi; i++) {
      if (i + 16 < hi) {
        __builtin_prefetch(items[i + 16]);
        __builtin_prefetch((const char*)items[i + 16] + 64);
      }
      PyObject* o = items[i];
      if (item >= 0) {   // elements are tuples (the reference's Polynomial constructor passes (coeff, [(0, 0)]) terms): take [item]
        if (!PyTuple_Check(o) || PyTuple_GET_SIZE(o) <= item) {
          Py_ssize_t expect = -1;
          bad_type.compare_exchange_strong(expect, i);
          return;
        }
        o = PyTuple_GET_ITEM(o, item);
      }
      if (!PyLong_Check(o)) {   // (bool and int subclasses are PyLong too; anything else is a TypeError like pyo3's extraction)
        Py_ssize_t expect = -1;
        bad_type.compare_exchange_strong(expect, i);
        return;
      }
      if (!long_to_limbs(o, out + (size_t)i * nlimbs, nlimbs)) slow[t].push_back(i);
    }
  };
  auto element = [&](Py_ssize_t i) { return item >= 0 ? PyTuple_GET_ITEM(items[i], item) : items[i]; };
  if (nt == 1) {
    work(0);
  } else {
    // the GIL stays with this thread: the workers only READ immutable int objects that the (unchangeable, we hold the GIL)
    // sequence keeps alive
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (int t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
  }
  if (bad_type.load() >= 0) {
    const Py_ssize_t i = bad_type.load();
    PyErr_Format(PyExc_TypeError, "element %zd: '%.100s' object cannot be interpreted as an integer%s", i,
                 Py_TYPE(items[i])->tp_name, item >= 0 ? " term" : "");
    Py_DECREF(fast);
    return nullptr;
  }
  // slow path: negative or oversized values -> v mod modulus (Fr::from(BigUint) semantics; negatives as Python's %)
  for (int t = 0; t < nt; t++) {
    for (Py_ssize_t i : slow[t]) {
      if (modulus == Py_None) {
        Py_DECREF(fast);
        PyErr_Format(PyExc_OverflowError, "element %zd is negative or does not fit %d 64-bit limbs", i, nlimbs);
        return nullptr;
      }
      if (!allow_negative && view_long(element(i)).negative) {   // BigUint extraction fails in the reference's pyo3 layer
        Py_DECREF(fast);
        PyErr_SetString(PyExc_OverflowError, "can't convert negative int to unsigned");
        return nullptr;
      }
      PyObject* red = PyNumber_Remainder(element(i), modulus);
      if (!red) {
        Py_DECREF(fast);
        return nullptr;
      }
      const bool ok = PyLong_Check(red) && long_to_limbs(red, out + (size_t)i * nlimbs, nlimbs);
      Py_DECREF(red);
      if (!ok) {
        Py_DECREF(fast);
        PyErr_Format(PyExc_OverflowError, "element %zd: reduced value does not fit %d limbs", i, nlimbs);
        return nullptr;
      }
    }
  }
  Py_DECREF(fast);
  return PyLong_FromSsize_t(n);
}

// ---- limbs_to_ints -----------------------------------------------------------------------------------------------------------
// Object creation needs the GIL, so this direction is serial; it builds the PyLong digit array directly (no byte-array detour).
inline PyObject* limbs_to_long(const uint64_t* in, int nlimbs) {
  int top = nlimbs - 1;
  while (top >= 0 && in[top] == 0) top--;
  if (top < 0) return PyLong_FromLong(0);
  if (top == 0) return PyLong_FromUnsignedLongLong(in[0]);
  const int bits = top * 64 + (64 - __builtin_clzll(in[top]));
  const int nd = (bits + 29) / 30;
#if PY_VERSION_HEX >= 0x030C0000
  PyLongObject* v = _PyLong_New(nd);
  if (!v) return nullptr;
  digit* d = v->long_value.ob_digit;
#else
  PyLongObject* v = _PyLong_New(nd);
  if (!v) return nullptr;
  digit* d = v->ob_digit;
#endif
  for (int i = 0; i < nd; i++) {
    const int pos = i * 30, limb = pos >> 6, off = pos & 63;
    uint64_t x = in[limb] >> off;
    if (off > 34 && limb + 1 <= top) x |= in[limb + 1] << (64 - off);
    d[i] = (digit)(x & 0x3fffffffu);
  }
  return (PyObject*)v;   // _PyLong_New sets a positive sign and the digit count
}

PyObject* limbs_to_ints(PyObject*, PyObject* args) {
  unsigned long long addr;
  Py_ssize_t count;
  int nlimbs;
  if (!PyArg_ParseTuple(args, "Kni", &addr, &count, &nlimbs)) return nullptr;
  if (nlimbs < 1 || nlimbs > 16 || count < 0) {
    PyErr_SetString(PyExc_ValueError, "bad count / nlimbs");
    return nullptr;
  }
  const uint64_t* in = (const uint64_t*)(uintptr_t)addr;
  PyObject* list = PyList_New(count);
  if (!list) return nullptr;
  for (Py_ssize_t i = 0; i < count; i++) {
    PyObject* v = limbs_to_long(in + (size_t)i * nlimbs, nlimbs);
    if (!v) {
      Py_DECREF(list);
      return nullptr;
    }
    PyList_SET_ITEM(list, i, v);
  }
  return list;
}

PyObject* set_threads(PyObject*, PyObject* args) {
  int n;
  if (!PyArg_ParseTuple(args, "i", &n)) return nullptr;
  g_threads = n;
  Py_RETURN_NONE;
}
PyObject* get_threads(PyObject*, PyObject*) { return PyLong_FromLong(thread_count((size_t)1 << 30)); }

PyMethodDef methods[] = {
    {"ints_to_limbs", ints_to_limbs, METH_VARARGS,
     "ints_to_limbs(seq, addr, nlimbs, modulus=None, item=-1) -> n.  Writes len(seq)*nlimbs little-endian uint64 words at addr; values that "
     "are negative or wider than nlimbs words are reduced mod `modulus` (OverflowError without one)."},
    {"limbs_to_ints", limbs_to_ints, METH_VARARGS, "limbs_to_ints(addr, count, nlimbs) -> list[int]"},
    {"set_threads", set_threads, METH_VARARGS, "set_threads(n): host threads used by ints_to_limbs (0 = automatic)"},
    {"get_threads", get_threads, METH_NOARGS, "host threads a large conversion uses"},
    {nullptr, nullptr, 0, nullptr}};

PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_marshal", "list[int] <-> limb arrays for the zksnake_b200 binding layer", -1, methods,
                      nullptr, nullptr, nullptr, nullptr};

}  // namespace

PyMODINIT_FUNC PyInit__marshal(void) { return PyModule_Create(&moddef); }
