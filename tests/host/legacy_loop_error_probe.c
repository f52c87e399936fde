#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#define NPY_TARGET_VERSION NPY_2_0_API_VERSION
#include <Python.h>
#include <numpy/arrayobject.h>
#include <numpy/ufuncobject.h>
static PyUFuncGenericFunction old_loop;
static int mode = 0;
static void my_loop(char** args, const npy_intp* dims, const npy_intp* steps, void* data) {
    if (mode == 1) {
        PyGILState_STATE st = PyGILState_Ensure();
        PyErr_SetString(PyExc_RuntimeError, "injected failure inside a legacy loop");
        PyGILState_Release(st);
        return;
    }
    old_loop(args, dims, steps, data);
}
static PyObject* install(PyObject* self, PyObject* args) {
    PyObject* np = PyImport_ImportModule("numpy");
    PyObject* uf = PyObject_GetAttrString(np, "add");
    int sig[3] = {NPY_FLOAT, NPY_FLOAT, NPY_FLOAT};
    int r = PyUFunc_ReplaceLoopBySignature((PyUFuncObject*)uf, my_loop, sig, &old_loop);
    return PyLong_FromLong(r);
}
static PyObject* setmode(PyObject* self, PyObject* args) { if (!PyArg_ParseTuple(args, "i", &mode)) return NULL; Py_RETURN_NONE; }
static PyMethodDef m[] = {{"install", install, METH_NOARGS, ""}, {"setmode", setmode, METH_VARARGS, ""}, {0}};
static struct PyModuleDef d = {PyModuleDef_HEAD_INIT, "probe", 0, -1, m};
PyMODINIT_FUNC PyInit_probe(void) { PyObject* mod = PyModule_Create(&d); import_array(); import_umath(); return mod; }
