"""Gram-predicted BatchNorm statistics and folded BatchNorm backward, restated in fp64.  TEST INFRASTRUCTURE ONLY
(same rules as pointnet_oracle.py: only tests/, smoke() and bench.py's CPU legs may import it).

The CUDA training step never materialises the pre-BN output `y` of conv5 nor of global_feat, nor the gradient `dy` with
respect to them.  It relies on three identities of  y = a W^T + b  followed by train-mode BatchNorm
(pcs.py:110, 113: `F.relu(self.bn5(self.conv5(x)))`, `F.relu(self.bn_global(self.global_feat(x)))`), restated here on the
tensors of `pointnet_oracle.forward_train` so that tests/test_folded_bn_cpu.py can show that they reproduce the
reference's autograd result (pcs.py:254) exactly:

 (1) batch statistics of y from the Gram matrix of its INPUT:
        sum_p y      = W s + n b                       s = sum_p a,  G = a^T a
        sum_p y^2    = diag(W G W^T) + 2 b (W s) + n b^2
 (2) BatchNorm backward is affine in (dz, y):  dy = A dz + Bc y + Cc  (per-channel A, Bc, Cc), hence
        dW = dy^T a = diag(A) Q + diag(Bc) (W G + b s^T) + Cc s^T          Q = dz^T a   (the raw weight-gradient GEMM)
        sum_p dz y  = rowdot(Q, W) + b sum_p dz                            (-> dgamma without y)
 (3) the data gradient needs neither y nor dy:
        dy W = dz (diag(A) W) + a (W^T diag(Bc) W) + 1 (Bc*b + Cc)^T W
     i.e. ONE GEMM over the concatenated operand [dz | a].
For global_feat dz is the max-pool gradient: one non-zero per (cloud, channel) at the arg-max row (pcs.py:114), so Q and
the first term of (3) are sparse row gathers / scatters.
"""
import numpy as np

BN_EPS = 1e-5


def gram(a):
    """s = sum_p a[p, :],  G = a^T a  (what the MN-major tcgen05 GEMM accumulates with A = B = a)."""
    a = np.asarray(a, np.float64)
    return a.sum(0), a.T @ a


def predicted_stats(W, b, s, G, n):
    """identity (1): {sum y, sum y^2} of y = a W^T + b over n rows, from s and G only."""
    Ws = W @ s
    s1 = Ws + n * b
    s2 = np.einsum("ck,kj,cj->c", W, G, W) + 2.0 * b * Ws + n * b * b
    return s1, s2


def bn_from_sums(s1, s2, n, gamma, beta):
    mean = s1 / n
    var = np.maximum(s2 / n - mean * mean, 0.0)
    invstd = 1.0 / np.sqrt(var + BN_EPS)
    return mean, invstd, gamma * invstd, beta - mean * gamma * invstd


def bwd_coefficients(sum_dz, sum_dz_y, mean, invstd, gamma, n):
    """dy = A dz + Bc y + Cc;  also returns dgamma = sum dz*yhat."""
    dgamma = invstd * (sum_dz_y - mean * sum_dz)
    A = gamma * invstd
    Bc = -A * invstd * dgamma / n
    Cc = -A * (sum_dz / n) - Bc * mean
    return A, Bc, Cc, dgamma


def folded_layer_backward(W, b, s, G, n, gamma, mean, invstd, Q, sum_dz):
    """identities (2) and (3) for one layer.  Returns dW, db, dgamma, dbeta and the operands of the data-gradient GEMM:
    W_dz = diag(A) W  (multiplies dz),  S = W^T diag(Bc) W  (multiplies a),  const row."""
    sum_dz_y = np.einsum("ck,ck->c", Q, W) + b * sum_dz
    A, Bc, Cc, dgamma = bwd_coefficients(sum_dz, sum_dz_y, mean, invstd, gamma, n)
    dW = A[:, None] * Q + Bc[:, None] * (W @ G + np.outer(b, s)) + np.outer(Cc, s)
    db = A * sum_dz + Bc * (W @ s + n * b) + n * Cc              # zero in exact arithmetic
    W_dz = A[:, None] * W
    S = W.T @ (Bc[:, None] * W)
    const = (Bc * b + Cc) @ W
    return dict(dW=dW, db=db, dgamma=dgamma, dbeta=sum_dz, W_dz=W_dz, S=S, const=const)


def trunk_tail_backward(cache, dg):
    """Backward of  conv5 -> bn5 -> relu -> global_feat -> bn_global -> relu -> max  given dg = d loss / d pooled feature
    (B, 1024), using only: a3 (input of conv5), a4 (input of global_feat), the arg-max rows and the pooled pre-BN values.
    Returns the parameter gradients of both layers and da3 (gradient wrt conv5's input)."""
    B, N = cache["B"], cache["N"]
    n = B * N
    L = cache["layers"]
    out = {}
    # ---- global_feat (sparse dz: one entry per cloud and channel)
    c5 = L["global_feat"]
    W5 = c5["W"]
    a4 = c5["a_in"]
    b5 = c5["b"]
    yhat5, invstd5, gamma5 = c5["bn"]
    arg = cache["argmax"]                                              # (B, 1024) row within the cloud
    rows = arg + (np.arange(B) * N)[:, None]
    s4, G4 = gram(a4)
    s1p, s2p = predicted_stats(W5, b5, s4, G4, n)
    mean5, invstd5p, _, _ = bn_from_sums(s1p, s2p, n, gamma5, 0.0)
    relu_on = np.take_along_axis(c5["relu"].reshape(B, N, -1), arg[:, None, :], 1)[:, 0, :]
    dzv = dg * relu_on                                                 # (B, 1024)
    Q5 = np.zeros_like(W5)
    for b in range(B):
        Q5 += dzv[b][:, None] * a4[rows[b]]                            # row gather
    f5 = folded_layer_backward(W5, b5, s4, G4, n, gamma5, mean5, invstd5p, Q5, dzv.sum(0))
    out["global_feat.weight"], out["global_feat.bias"] = f5["dW"][:, :, None], f5["db"]
    out["bn_global.weight"], out["bn_global.bias"] = f5["dgamma"], f5["dbeta"]
    da4 = a4 @ f5["S"] + f5["const"]
    for b in range(B):
        np.add.at(da4, rows[b], dzv[b][:, None] * f5["W_dz"])          # sparse rows
    # ---- conv5
    c4 = L["conv5"]
    W4 = c4["W"]
    a3 = c4["a_in"]
    b4 = c4["b"]
    _, invstd4, gamma4 = c4["bn"]
    dz4 = da4 * (a4 > 0)                                               # the mask comes from the stored activation
    s3, G3 = gram(a3)
    s1p, s2p = predicted_stats(W4, b4, s3, G3, n)
    mean4, invstd4p, _, _ = bn_from_sums(s1p, s2p, n, gamma4, 0.0)
    Q4 = dz4.T @ a3
    f4 = folded_layer_backward(W4, b4, s3, G3, n, gamma4, mean4, invstd4p, Q4, dz4.sum(0))
    out["conv5.weight"], out["conv5.bias"] = f4["dW"][:, :, None], f4["db"]
    out["bn5.weight"], out["bn5.bias"] = f4["dgamma"], f4["dbeta"]
    da3 = np.concatenate([dz4, a3], axis=1) @ np.concatenate([f4["W_dz"], f4["S"]], axis=0) + f4["const"]
    out["_da3"] = da3
    out["_pred"] = dict(invstd5=invstd5p, invstd4=invstd4p, invstd5_ref=invstd5, invstd4_ref=invstd4)
    return out
