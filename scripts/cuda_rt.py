"""Minimal ctypes binding to libcudart for the bring-up / sweep scripts (no torch import: fast subprocess start)."""
import ctypes as C
import glob
import numpy as np

_rt = None


def rt():
    global _rt
    if _rt is None:
        cands = sorted(glob.glob("/usr/local/cuda/lib64/libcudart.so*"))
        _rt = C.CDLL(cands[-1])
        _rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        _rt.cudaFree.argtypes = [C.c_void_p]
        _rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        _rt.cudaMemset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
        _rt.cudaEventCreate.argtypes = [C.POINTER(C.c_void_p)]
        _rt.cudaEventRecord.argtypes = [C.c_void_p, C.c_void_p]
        _rt.cudaEventSynchronize.argtypes = [C.c_void_p]
        _rt.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
        _rt.cudaGetErrorString.restype = C.c_char_p
        _rt.cudaGetErrorString.argtypes = [C.c_int]
        _rt.cudaMallocHost.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    return _rt


def check(code):
    if code != 0:
        raise RuntimeError(f"CUDA error {code}: {rt().cudaGetErrorString(code).decode()}")


class DevBuf:
    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        self.nbytes = nbytes
        check(rt().cudaMalloc(C.byref(self.ptr), nbytes))

    @classmethod
    def from_numpy(cls, a):
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes)
        check(rt().cudaMemcpy(b.ptr, a.ctypes.data, a.nbytes, 1))
        return b

    def upload(self, a):
        a = np.ascontiguousarray(a)
        check(rt().cudaMemcpy(self.ptr, a.ctypes.data, a.nbytes, 1))

    def to_numpy(self, dtype, count):
        out = np.empty(count, dtype)
        check(rt().cudaMemcpy(out.ctypes.data, self.ptr, out.nbytes, 2))
        return out

    def zero(self):
        check(rt().cudaMemset(self.ptr, 0, self.nbytes))

    def data_ptr(self):
        return self.ptr.value

    def free(self):
        if self.ptr:
            rt().cudaFree(self.ptr)
            self.ptr = C.c_void_p()


def sync():
    check(rt().cudaDeviceSynchronize())


class Timer:
    def __init__(self):
        self.a, self.b = C.c_void_p(), C.c_void_p()
        check(rt().cudaEventCreate(C.byref(self.a)))
        check(rt().cudaEventCreate(C.byref(self.b)))

    def start(self):
        check(rt().cudaEventRecord(self.a, None))

    def stop(self):
        check(rt().cudaEventRecord(self.b, None))
        check(rt().cudaEventSynchronize(self.b))
        ms = C.c_float()
        check(rt().cudaEventElapsedTime(C.byref(ms), self.a, self.b))
        return ms.value
