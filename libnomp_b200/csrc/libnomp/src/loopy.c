/* Bridge between the C runtime and the embedded Python side (nomp_bridge + user transform scripts).
 *
 * Same role and call sequence as the reference's bridge (reference src/loopy.c:31-500): initialise/borrow the
 * interpreter, turn the kernel string into a kernel object, run the user's transform / annotate functions on it,
 * realise the reduce clause, fix JIT parameters, and pull out name, device source and launch-size expressions.
 * What differs: the Python modules behind it (nomp_bridge instead of loopy_api/reduction, no loopy, libclang or
 * SymEngine), and every entry point takes the GIL itself so the library can also live inside a Python process.
 * The file keeps the name src/loopy.c because the reference tests match it in error strings
 * (reference tests/nomp-api-100-impl.h:45-46).
 */
#include "nomp-aux.h"
#include "nomp-impl.h"
#include "nomp-loopy.h"

static char backend_name[NOMP_MAX_BUFFER_SIZE + 1];
static int interpreter_owned = 0; /* we called Py_InitializeEx, so we may call Py_FinalizeEx */

#define BRIDGE_MODULE "nomp_bridge"

/* Fail with a NOMP_PY_CALL_FAILURE-style log entry if `obj` is NULL/false. */
#define py_require(obj, errno_, ...)                                                                             \
  do {                                                                                                           \
    if (!(obj)) {                                                                                                \
      if (PyErr_Occurred()) {                                                                                    \
        if (nomp_log_get_verbose() >= NOMP_INFO) PyErr_Print();                                                  \
        PyErr_Clear();                                                                                           \
      }                                                                                                          \
      return nomp_log((errno_), NOMP_ERROR, __VA_ARGS__);                                                        \
    }                                                                                                            \
  } while (0)

/* Run `expr` (an int-returning call) with the GIL held. */
#define WITH_GIL(expr)                                                                                           \
  do {                                                                                                           \
    PyGILState_STATE gil_ = PyGILState_Ensure();                                                                 \
    int err_ = (expr);                                                                                           \
    PyGILState_Release(gil_);                                                                                    \
    return err_;                                                                                                 \
  } while (0)

/* ------------------------------------------------------------------------------------------------------------ */
static int append_to_sys_path(const char *path) {
  if (path == NULL || path[0] == '\0') return 0;
  PyObject *sys_path = PySys_GetObject("path"); /* borrowed */
  py_require(sys_path, NOMP_PY_CALL_FAILURE, "Getting attribute sys.path failed.");
  PyObject *entry = PyUnicode_FromString(path);
  py_require(entry, NOMP_PY_CALL_FAILURE, "Converting C string \"%s\" to python string failed.", path);
  int present = PySequence_Contains(sys_path, entry);
  int rc = present == 1 ? 0 : PyList_Append(sys_path, entry);
  Py_DECREF(entry);
  py_require(rc == 0, NOMP_PY_CALL_FAILURE, "Appending path \"%s\" to the sys.path failed.", path);
  return 0;
}

int nomp_py_append_to_sys_path(const char *path) { WITH_GIL(append_to_sys_path(path)); }

static int py_init_locked(const nomp_config_t *cfg) {
  nomp_check(append_to_sys_path("."));
  char *py_dir = nomp_str_cat(2, PATH_MAX, cfg->install_dir, "/python");
  int err = append_to_sys_path(py_dir);
  free(py_dir);
  if (err) return err;
  nomp_check(append_to_sys_path(cfg->scripts_dir));
  return 0;
}

int nomp_py_init(const nomp_config_t *cfg) {
  strncpy(backend_name, cfg->backend, NOMP_MAX_BUFFER_SIZE);
  backend_name[NOMP_MAX_BUFFER_SIZE] = '\0';
  if (!Py_IsInitialized()) {
    Py_InitializeEx(0); /* no signal handlers: we are a library (reference src/loopy.c:38) */
    interpreter_owned = 1;
  }
  WITH_GIL(py_init_locked(cfg));
}

int nomp_py_finalize(int interpreter) {
  if (interpreter && interpreter_owned && Py_IsInitialized()) {
    interpreter_owned = 0;
    if (Py_FinalizeEx() < 0) return nomp_log(NOMP_PY_CALL_FAILURE, NOMP_ERROR, "Finalizing the Python interpreter failed.");
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
static int check_module_locked(const char *module, const char *function) {
  PyObject *py_module = PyImport_ImportModule(module);
  py_require(py_module, NOMP_PY_CALL_FAILURE, "Importing Python module \"%s\" failed.", module);
  PyObject *py_function = PyObject_GetAttrString(py_module, function);
  Py_DECREF(py_module);
  py_require(py_function, NOMP_PY_CALL_FAILURE, "Importing Python function \"%s\" from module \"%s\" failed.",
             function, module);
  Py_DECREF(py_function);
  return 0;
}

int nomp_py_check_module(const char *module, const char *function) {
  if (module == NULL || function == NULL) {
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Module name and/or function name not provided.");
  }
  WITH_GIL(check_module_locked(module, function));
}

/* Call nomp_bridge.<name>(args...); returns a new reference or NULL with the Python error set. */
static PyObject *call_bridge(const char *name, const char *fmt, ...) {
  PyObject *module = PyImport_ImportModule(BRIDGE_MODULE);
  if (!module) return NULL;
  PyObject *fn = PyObject_GetAttrString(module, name);
  Py_DECREF(module);
  if (!fn) return NULL;
  va_list ap;
  va_start(ap, fmt);
  PyObject *args = Py_VaBuildValue(fmt, ap);
  va_end(ap);
  PyObject *result = NULL;
  if (args) {
    result = PyObject_CallObject(fn, args);
    Py_DECREF(args);
  }
  Py_DECREF(fn);
  return result;
}

static int c_to_loopy_locked(PyObject **kernel, const char *src) {
  PyObject *knl = call_bridge("c_to_loopy", "(ss)", src, backend_name);
  py_require(knl, NOMP_LOOPY_CONVERSION_FAILURE, "Converting C source to loopy kernel failed.");
  *kernel = knl;
  return 0;
}

int nomp_py_c_to_loopy(PyObject **kernel, const char *src) { WITH_GIL(c_to_loopy_locked(kernel, src)); }

static int replace_kernel(PyObject **kernel, PyObject *result) {
  Py_DECREF(*kernel);
  *kernel = result;
  return 0;
}

static int transform_locked(PyObject **kernel, const char *file, const char *function, const PyObject *context) {
  PyObject *module = PyImport_ImportModule(file);
  py_require(module, NOMP_PY_CALL_FAILURE, "Importing Python module: \"%s\" failed.", file);
  PyObject *fn = PyObject_GetAttrString(module, function);
  Py_DECREF(module);
  py_require(fn, NOMP_PY_CALL_FAILURE, "Importing Python function \"%s\" from  module \"%s\" failed.", function, file);
  if (!PyCallable_Check(fn)) {
    Py_DECREF(fn);
    return nomp_log(NOMP_PY_CALL_FAILURE, NOMP_ERROR, "Python function \"%s\" from  module \"%s\" is not callable.",
                    function, file);
  }
  PyObject *result = PyObject_CallFunctionObjArgs(fn, *kernel, (PyObject *)context, NULL);
  Py_DECREF(fn);
  py_require(result, NOMP_PY_CALL_FAILURE, "Calling Python function \"%s\" from module \"%s\" failed.", function, file);
  return replace_kernel(kernel, result);
}

int nomp_py_transform(PyObject **kernel, const char *file, const char *function, const PyObject *context) {
  if (!kernel || !*kernel) return 0;
  WITH_GIL(transform_locked(kernel, file, function, context));
}

static int set_annotate_locked(PyObject **annotate, const char *file) {
  PyObject *module = PyImport_ImportModule(file);
  py_require(module, NOMP_PY_CALL_FAILURE, "Importing python module \"%s\" failed.", file);
  PyObject *fn = PyObject_GetAttrString(module, "annotate");
  Py_DECREF(module);
  py_require(fn, NOMP_PY_CALL_FAILURE, "Failed to find annotate function in file \"%s\".", file);
  if (!PyCallable_Check(fn)) {
    Py_DECREF(fn);
    return nomp_log(NOMP_PY_CALL_FAILURE, NOMP_ERROR, "Annotate function is not callable.");
  }
  Py_XDECREF(*annotate);
  *annotate = fn;
  return 0;
}

int nomp_py_set_annotate_func(PyObject **annotate, const char *file) {
  if (file == NULL || file[0] == '\0') return 0; /* no annotations script configured */
  WITH_GIL(set_annotate_locked(annotate, file));
}

static int annotate_locked(PyObject **kernel, PyObject *function, const PyObject *annotations,
                           const PyObject *context) {
  PyObject *result;
  if (function) {
    result = PyObject_CallFunctionObjArgs(function, *kernel, (PyObject *)annotations, (PyObject *)context, NULL);
  } else {
    result = call_bridge("annotate_passthrough", "(OOO)", *kernel, (PyObject *)annotations, (PyObject *)context);
  }
  py_require(result, NOMP_PY_CALL_FAILURE, "Annotating loopy kernel failed.");
  return replace_kernel(kernel, result);
}

int nomp_py_annotate(PyObject **kernel, PyObject *function, const PyObject *annotations, const PyObject *context) {
  if (!kernel || !*kernel) return 0;
  WITH_GIL(annotate_locked(kernel, function, annotations, context));
}

static int realize_reduction_locked(PyObject **kernel, const char *var, const char *op, const PyObject *context) {
  PyObject *result = call_bridge("realize_reduction", "(OssO)", *kernel, var, op, (PyObject *)context);
  py_require(result, NOMP_PY_CALL_FAILURE, "Calling realize_reduction() function failed.");
  return replace_kernel(kernel, result);
}

int nomp_py_realize_reduction(PyObject **kernel, const char *var, const char *op, const PyObject *context) {
  WITH_GIL(realize_reduction_locked(kernel, var, op, context));
}

static int fix_parameters_locked(PyObject **kernel, const PyObject *dict) {
  PyObject *result = call_bridge("fix_parameters", "(OO)", *kernel, (PyObject *)dict);
  py_require(result, NOMP_PY_CALL_FAILURE, "Calling loopy.fix_parameters() failed.");
  return replace_kernel(kernel, result);
}

int nomp_py_fix_parameters(PyObject **kernel, const PyObject *dict) { WITH_GIL(fix_parameters_locked(kernel, dict)); }

static int name_and_src_locked(char **name, char **src, const PyObject *kernel, const PyObject *context) {
  PyObject *py_name = call_bridge("get_knl_name", "(O)", (PyObject *)kernel);
  py_require(py_name, NOMP_LOOPY_KNL_NAME_NOT_FOUND, "Unable to get loopy kernel name.");
  const char *name_ = PyUnicode_AsUTF8(py_name);
  *name = name_ ? strndup(name_, NOMP_MAX_BUFFER_SIZE) : NULL;
  Py_DECREF(py_name);
  py_require(*name, NOMP_LOOPY_KNL_NAME_NOT_FOUND, "Unable to get loopy kernel name.");

  PyObject *py_src = call_bridge("get_knl_src", "(OO)", (PyObject *)kernel, (PyObject *)context);
  if (!py_src) {
    char who[NOMP_MAX_BUFFER_SIZE + 1];
    strncpy(who, *name, NOMP_MAX_BUFFER_SIZE), who[NOMP_MAX_BUFFER_SIZE] = '\0';
    free(*name), *name = NULL;
    py_require(py_src, NOMP_LOOPY_CODEGEN_FAILURE, "Backend code generation from loopy kernel \"%s\" failed.", who);
  }
  const char *src_ = PyUnicode_AsUTF8(py_src);
  *src = src_ ? strdup(src_) : NULL;
  Py_DECREF(py_src);
  py_require(*src, NOMP_LOOPY_CODEGEN_FAILURE, "Backend code generation from loopy kernel \"%s\" failed.", *name);
  return 0;
}

int nomp_py_get_knl_name_and_src(char **name, char **src, const PyObject *kernel, const PyObject *context) {
  WITH_GIL(name_and_src_locked(name, src, kernel, context));
}

static int grid_size_locked(nomp_prog_t *prg, PyObject *kernel, const PyObject *context) {
  PyObject *sizes = call_bridge("get_grid_size", "(OO)", kernel, (PyObject *)context);
  py_require(sizes, NOMP_LOOPY_GRIDSIZE_FAILURE, "Unable to evaluate grid sizes from loopy kernel.");
  int ok = PyTuple_Check(sizes) && PyTuple_Size(sizes) == 2;
  for (int which = 0; ok && which < 2; which++) {
    PyObject *t = PyTuple_GetItem(sizes, which); /* borrowed */
    ok = PyTuple_Check(t) && PyTuple_Size(t) == 3;
    for (int d = 0; ok && d < 3; d++) {
      const char *s = PyUnicode_AsUTF8(PyTuple_GetItem(t, d));
      if (!s) {
        ok = 0;
        break;
      }
      char **slot = which == 0 ? &prg->sym_global[d] : &prg->sym_local[d];
      free(*slot);
      *slot = strdup(s);
    }
  }
  Py_DECREF(sizes);
  py_require(ok, NOMP_LOOPY_GRIDSIZE_FAILURE, "Grid size is not a pair of 3-tuples of strings.");
  prg->ndim = 3;
  return 0;
}

int nomp_py_get_grid_size(nomp_prog_t *prg, PyObject *kernel, const PyObject *context) {
  WITH_GIL(grid_size_locked(prg, kernel, context));
}

/* ---- helpers of the on-disk JIT cache: no error is logged, a failure only means "do not cache" ---------------- */
static int repr_locked(char **out, const PyObject *obj) {
  PyObject *r = obj ? PyObject_Repr((PyObject *)obj) : NULL;
  const char *s = r ? PyUnicode_AsUTF8(r) : NULL;
  *out = s ? strdup(s) : NULL;
  Py_XDECREF(r);
  PyErr_Clear();
  return *out == NULL;
}

int nomp_py_repr(char **out, const PyObject *obj) { WITH_GIL(repr_locked(out, obj)); }

static int module_file_locked(char **out, const char *module) {
  *out = NULL;
  PyObject *util = PyImport_ImportModule("importlib.util");
  PyObject *spec = util ? PyObject_CallMethod(util, "find_spec", "s", module) : NULL;
  PyObject *origin = (spec && spec != Py_None) ? PyObject_GetAttrString(spec, "origin") : NULL;
  const char *s = (origin && PyUnicode_Check(origin)) ? PyUnicode_AsUTF8(origin) : NULL;
  if (s) *out = strdup(s);
  Py_XDECREF(origin), Py_XDECREF(spec), Py_XDECREF(util);
  PyErr_Clear();
  return *out == NULL;
}

int nomp_py_module_file(char **out, const char *module) { WITH_GIL(module_file_locked(out, module)); }

/* ------------------------------------------------------------------------------------------------------------ */
PyObject *nomp_py_dict_new(void) {
  PyGILState_STATE gil = PyGILState_Ensure();
  PyObject *d = PyDict_New();
  PyGILState_Release(gil);
  return d;
}

static void dict_set(PyObject *dict, const char *key, PyObject *value) {
  if (dict && value) PyDict_SetItemString(dict, key, value);
  Py_XDECREF(value);
}

void nomp_py_dict_set_str(PyObject *dict, const char *key, const char *value) {
  PyGILState_STATE gil = PyGILState_Ensure();
  dict_set(dict, key, PyUnicode_FromString(value));
  PyGILState_Release(gil);
}

void nomp_py_dict_set_long(PyObject *dict, const char *key, long value) {
  PyGILState_STATE gil = PyGILState_Ensure();
  dict_set(dict, key, PyLong_FromLong(value));
  PyGILState_Release(gil);
}

void nomp_py_dict_set_double(PyObject *dict, const char *key, double value) {
  PyGILState_STATE gil = PyGILState_Ensure();
  dict_set(dict, key, PyFloat_FromDouble(value));
  PyGILState_Release(gil);
}

long nomp_py_dict_size(PyObject *dict) {
  if (!dict) return 0;
  PyGILState_STATE gil = PyGILState_Ensure();
  long n = (long)PyDict_Size(dict);
  PyGILState_Release(gil);
  return n;
}

void nomp_py_decref(PyObject **obj) {
  if (!obj || !*obj) return;
  if (Py_IsInitialized()) {
    PyGILState_STATE gil = PyGILState_Ensure();
    Py_DECREF(*obj);
    PyGILState_Release(gil);
  }
  *obj = NULL;
}
