/* C side of the jit bridge into the embedded interpreter (reference include/nomp-loopy.h:16-45).  Every function
 * acquires the GIL itself, so libnomp.so also works when it is loaded into a running Python process via ctypes. */
#ifndef LIBNOMP_B200_LOOPY_H_
#define LIBNOMP_B200_LOOPY_H_

#include "nomp-impl.h"

#ifdef __cplusplus
extern "C" {
#endif

int nomp_py_init(const nomp_config_t *cfg);
int nomp_py_append_to_sys_path(const char *path);
int nomp_py_check_module(const char *module, const char *function);
int nomp_py_c_to_loopy(PyObject **kernel, const char *src);
int nomp_py_transform(PyObject **kernel, const char *file, const char *function, const PyObject *context);
int nomp_py_set_annotate_func(PyObject **annotate, const char *file);
int nomp_py_annotate(PyObject **kernel, PyObject *function, const PyObject *annotations, const PyObject *context);
int nomp_py_realize_reduction(PyObject **kernel, const char *var, const char *op, const PyObject *context);
int nomp_py_fix_parameters(PyObject **kernel, const PyObject *dict);
int nomp_py_get_knl_name_and_src(char **name, char **src, const PyObject *kernel, const PyObject *context);
int nomp_py_get_grid_size(nomp_prog_t *prg, PyObject *kernel, const PyObject *context);
int nomp_py_finalize(int interpreter);
/* for the JIT cache key: repr() of a dict, and the source file behind an importable module (both malloc'ed) */
int nomp_py_repr(char **out, const PyObject *obj);
int nomp_py_module_file(char **out, const char *module);

/* small dict helpers so that nomp.c and the backend need no Python API knowledge of their own */
PyObject *nomp_py_dict_new(void);
void nomp_py_dict_set_str(PyObject *dict, const char *key, const char *value);
void nomp_py_dict_set_long(PyObject *dict, const char *key, long value);
void nomp_py_dict_set_double(PyObject *dict, const char *key, double value);
long nomp_py_dict_size(PyObject *dict);
void nomp_py_decref(PyObject **obj);

#ifdef __cplusplus
}
#endif

#endif
