/* Host-side glue of the drop-in API (no compute): KLT_Feature list <-> the structure-of-arrays the C ABI takes.
 * The reference walks its feature list in Python (trackFeatures.py:252-345); doing the same around a 20 us kernel costs
 * ~0.4 ms per 1000 features, so the two walks of KLTTrackFeatures live here.  Loaded with ctypes.PyDLL (GIL held);
 * a missing library only means the Python loops in trackFeatures.py run instead. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

static PyObject *s_x, *s_y, *s_val, *s_aff[3];

static int names(void) {
    if (s_x) return 0;
    s_x = PyUnicode_InternFromString("x");
    s_y = PyUnicode_InternFromString("y");
    s_val = PyUnicode_InternFromString("val");
    s_aff[0] = PyUnicode_InternFromString("aff_img");
    s_aff[1] = PyUnicode_InternFromString("aff_img_gradx");
    s_aff[2] = PyUnicode_InternFromString("aff_img_grady");
    return (s_x && s_y && s_val && s_aff[0] && s_aff[1] && s_aff[2]) ? 0 : -1;
}

static int as_double(PyObject *o, PyObject *name, double *out) {
    PyObject *v = PyObject_GetAttr(o, name);
    if (!v) return -1;
    *out = PyFloat_AsDouble(v);          /* float, int and NumPy scalars (selection leaves np.int32 coordinates, quirk Q13) */
    Py_DECREF(v);
    return (*out == -1.0 && PyErr_Occurred()) ? -1 : 0;
}

/* x, y, val of every feature; lost features (val < 0) travel as (-1, -1) like trackFeatures.py:253 leaves them */
int klt_featlist_gather(PyObject *list, Py_ssize_t n, double *x, double *y, int32_t *val) {
    if (names() || !PyList_Check(list) || PyList_GET_SIZE(list) != n) {
        if (!PyErr_Occurred()) PyErr_SetString(PyExc_TypeError, "feature list expected");
        return -1;
    }
    for (Py_ssize_t i = 0; i < n; i++) {
        PyObject *f = PyList_GET_ITEM(list, i);
        PyObject *v = PyObject_GetAttr(f, s_val);
        if (!v) return -1;
        PyObject *vi = PyNumber_Index(v);
        Py_DECREF(v);
        if (!vi) return -1;
        const long lv = PyLong_AsLong(vi);
        Py_DECREF(vi);
        if (lv == -1 && PyErr_Occurred()) return -1;
        val[i] = (int32_t)lv;
        if (lv >= 0) {
            if (as_double(f, s_x, &x[i]) || as_double(f, s_y, &y[i])) return -1;
        } else {
            x[i] = -1.0; y[i] = -1.0;
        }
    }
    return 0;
}

/* the write-back of KLTTrackFeatures (trackFeatures.py:253, 330-345): features that were live get the tracked position
 * and val 0, or (-1.0, -1.0, status) and lose their affine template */
int klt_featlist_scatter_tracked(PyObject *list, Py_ssize_t n, const double *x, const double *y, const int32_t *val,
                                 const int32_t *old_val) {
    if (names() || !PyList_Check(list) || PyList_GET_SIZE(list) != n) {
        if (!PyErr_Occurred()) PyErr_SetString(PyExc_TypeError, "feature list expected");
        return -1;
    }
    PyObject *zero = PyLong_FromLong(0), *minus1 = PyFloat_FromDouble(-1.0);
    int rc = (zero && minus1) ? 0 : -1;
    for (Py_ssize_t i = 0; i < n && !rc; i++) {
        if (old_val[i] < 0) continue;
        PyObject *f = PyList_GET_ITEM(list, i);
        if (val[i] == 0) {
            PyObject *fx = PyFloat_FromDouble(x[i]), *fy = PyFloat_FromDouble(y[i]);
            if (!fx || !fy || PyObject_SetAttr(f, s_x, fx) || PyObject_SetAttr(f, s_y, fy) || PyObject_SetAttr(f, s_val, zero)) rc = -1;
            Py_XDECREF(fx); Py_XDECREF(fy);
        } else {
            PyObject *v = PyLong_FromLong(val[i]);
            if (!v || PyObject_SetAttr(f, s_x, minus1) || PyObject_SetAttr(f, s_y, minus1) || PyObject_SetAttr(f, s_val, v)) rc = -1;
            Py_XDECREF(v);
            if (!rc) {
                const int has = PyObject_HasAttr(f, s_aff[0]);
                for (int k = 0; k < 3 && has && !rc; k++)
                    if (PyObject_SetAttr(f, s_aff[k], Py_None)) rc = -1;
            }
        }
    }
    Py_XDECREF(zero); Py_XDECREF(minus1);
    return rc;
}
