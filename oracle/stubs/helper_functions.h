/* Empty stand-in: see helper_cuda.h in this directory. */
