"""Generate tests/golden/*.json from the reference's OWN host code (oracle/_ref/libref_utils.so =
/root/reference/utils/utils.cu compiled unmodified).  Run in the build container, where
/root/reference exists; the JSON travels to the GPU box, /root/reference does not.

For END in {64,128,256,512,1024}: srand(10); A,B,C = generate_random_matrix(END); C=0 (sgemm.cu:12,52-56);
C = cpu_gemm(1, 0, B^T-buffer, A-buffer) which is the kernels' column-major NT result
(SURVEY.md section 8c oracle 1).  Stored: leading elements of A/B/C, selected C entries, sums, and a
sha256 of the raw float32 bytes of A, B and C.
"""
import hashlib, json, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import oracle as O

out_dir = Path(__file__).resolve().parents[1] / "tests" / "golden"
out_dir.mkdir(parents=True, exist_ok=True)
r = O.ref_utils()
assert r is not None, "needs oracle/_ref (run `make -C oracle`)"
gold = {}
for n in (64, 128, 256, 512, 1024):
    A, B, C = O.ref_make_inputs(n)
    # cpu_gemm is row-major square: Z = X*Y.  With X := B-buffer read row-major = B^T ... see SURVEY 8c:
    # kernels' C buffer == cpu_gemm(alpha, beta, B^T_buffer, A_buffer).  B^T_buffer: row-major (K x N)^T
    # Concretely Cbuf[m + n*M] = sum_k A[m+k*M]*B[n+k*N]; as row-major Z[i=n][j=m] = sum_k X[n][k] Y[k][m]
    # with X[n][k] = B[n + k*N] (i.e. X = transpose of the B buffer read row-major) and Y = A buffer row-major.
    X = np.ascontiguousarray(B.reshape(n, n).T)
    Z = np.zeros(n * n, np.float32)
    r.ref_cpu_gemm(1.0, 0.0, O._p(X.reshape(-1)), O._p(A), n, O._p(Z))
    gold[str(n)] = {
        "A_head": [float(x) for x in A[:8]], "B_head": [float(x) for x in B[:8]],
        "C_sel": {"0": float(Z[0]), "1": float(Z[1]), str(n): float(Z[n]), str(n * n - 1): float(Z[n * n - 1])},
        "C_sum": float(Z.astype(np.float64).sum()), "C_abs_sum": float(np.abs(Z.astype(np.float64)).sum()),
        "sha256": {"A": hashlib.sha256(A.tobytes()).hexdigest(), "B": hashlib.sha256(B.tobytes()).hexdigest(),
                   "C": hashlib.sha256(Z.tobytes()).hexdigest()},
    }
    print(n, gold[str(n)]["C_sel"], gold[str(n)]["C_sum"], gold[str(n)]["C_abs_sum"])
(out_dir / "ref_cpu_gemm.json").write_text(json.dumps(gold, indent=1))
