/* Empty stand-in: the reference includes NVIDIA cuda-samples' helper_cuda.h
 * (an un-vendored dependency, reference cuda-samples/ is empty) but uses nothing from it. */
