"""A second, independent restatement of the reference's tree interpreter in plain NumPy float32
(pure-Python loops, small cases only).  Follows genetic_programming.py:356-388 (body_fun/foriloop)
statement by statement; used to cross-check the C oracle."""
import numpy as np

from multitreegp_b200 import abi

F32 = np.float32


def node_functions(library):
    """node_function_list of genetic_programming.py:135-180 for the closed operator set."""
    fns = [lambda x, y, d: F32(0.0), lambda x, y, d: F32(0.0)]
    with np.errstate(all="ignore"):
        table = {
            abi.OP_ADD: lambda x, y, d: F32(x + y), abi.OP_SUB: lambda x, y, d: F32(x - y),
            abi.OP_MUL: lambda x, y, d: F32(x * y), abi.OP_DIV: lambda x, y, d: F32(np.divide(x, y)),
            abi.OP_MAX: lambda x, y, d: F32(np.maximum(x, y)), abi.OP_MIN: lambda x, y, d: F32(np.minimum(x, y)),
            abi.OP_SIN: lambda x, y, d: F32(np.sin(x)), abi.OP_COS: lambda x, y, d: F32(np.cos(x)),
            abi.OP_EXP: lambda x, y, d: F32(np.exp(x)), abi.OP_LOG: lambda x, y, d: F32(np.log(x)),
            abi.OP_SQRT: lambda x, y, d: F32(np.sqrt(x)), abi.OP_TANH: lambda x, y, d: F32(np.tanh(x)),
            abi.OP_NEG: lambda x, y, d: F32(-x), abi.OP_ABS: lambda x, y, d: F32(np.abs(x)),
            abi.OP_SQUARE: lambda x, y, d: F32(x * x),
        }
    for i in range(2, library.n_ids):
        opc, arg = int(library.opcodes[i]), int(library.args[i])
        if opc == abi.OP_VAR:
            fns.append(lambda x, y, d, j=arg: F32(d[j]))
        else:
            fns.append(table[opc])
    return fns


def foriloop(tree, data, fns):
    """x, _ = fori_loop(0, max_nodes, body_fun, (tree, data)); return x[-1, -1]"""
    tree = np.array(tree, dtype=F32, copy=True)
    data = np.asarray(data, dtype=F32)
    n = tree.shape[0]
    with np.errstate(all="ignore"):
        for i in range(n):
            f_idx, a_idx, b_idx, coefficient = tree[i]
            x = tree[int(a_idx), 3]      # negative indices wrap like NumPy/JAX
            y = tree[int(b_idx), 3]
            k = min(max(int(f_idx), 0), len(fns) - 1)  # lax.switch clamps
            value = coefficient if f_idx == 1 else fns[k](x, y, data)
            tree[i, 3] = value
    return tree[-1, -1]


# ---- a second restatement of the control evaluators (NumPy float32, Euler + ConstantStepSize) --------
# Written directly from the reference sources, independently of oracle/mtgp_oracle.c, to catch
# transcription errors in the C oracle.  Transcendentals come from NumPy, so agreement is to ~1e-5 over
# short horizons, not bitwise.
def acrobot_drift(state, u, p):
    """acrobot.py:51-72"""
    l1, l2, m1, m2 = [F32(v) for v in p]
    lc1, lc2, moi1, moi2, g = F32(0.5) * l1, F32(0.5) * l2, F32(1.0), F32(1.0), F32(9.81)
    control = np.clip(F32(u), F32(-1), F32(1))
    th1, th2, th1d, th2d = [F32(v) for v in state]
    pi2 = F32(np.pi / 2)
    d1 = m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * np.cos(th2)) + moi1 + moi2
    d2 = m2 * (lc2 ** 2 + l1 * lc2 * np.cos(th2)) + moi2
    phi2 = m2 * lc2 * g * np.cos(th1 + th2 - pi2)
    phi1 = -m2 * l1 * lc2 * th2d ** 2 * np.sin(th2) - 2 * m2 * l1 * lc2 * th1d * th2d * np.sin(th1) \
        + (m1 * lc1 + m2 * l1) * g * np.cos(th1 - pi2) + phi2
    th2acc = (control + d2 / d1 * phi1 - m2 * l1 * lc2 * th1d ** 2 * np.sin(th2) - phi2) / (m2 * lc2 ** 2 + moi2 - d2 ** 2 / d1)
    th1acc = -(d2 * th2acc + phi1) / d1
    return np.array([th1d, th2d, th1acc, th2acc], dtype=F32)


def acrobot_obs(x, noise):
    """control_environment_base.py:43-48 + acrobot.py:29-32 (C = I, W = obs_noise I folded into `noise`)"""
    out = (np.eye(4, dtype=F32) @ x.astype(F32) + noise).astype(F32)
    pi = F32(np.pi)
    out[0] = (out[0] + pi) % (2 * pi) - pi
    out[1] = (out[1] + pi) % (2 * pi) - pi
    return out.astype(F32)


def acrobot_fitness(xs, us, ts):
    """acrobot.py:77-84"""
    reached = -np.cos(xs[:, 0]) - np.cos(xs[:, 0] + xs[:, 1]) > 1.5
    first_success = int(np.argmax(reached))
    control_cost = (us * F32(0.01) * us).astype(F32)
    costs = np.where((ts / (ts[1] - ts[0])) > first_success, F32(0), control_cost)
    return F32(first_success + (first_success == 0) * ts.shape[0] + np.sum(costs, dtype=F32))


def dynamic_acrobot_rollout(candidate, fns, state_size, x0, ts, dt0, params, noise_at):
    """dynamic_evaluate.py:65-118 with diffrax Euler/ConstantStepSize/SaveAt(ts) (steps aligned with ts are
    not assumed: linear dense output).  noise_at(t) -> f32[4] is the observation-noise term at time t."""
    S = state_size
    state_eq, readout = candidate[:S], candidate[S:]
    with np.errstate(all="ignore"):
        def drift(t, xa):
            x, a = xa[:4], xa[4:]
            y = acrobot_obs(x, noise_at(t))
            u = F32(foriloop(readout[0], np.concatenate([np.zeros(4, F32), a, np.zeros(1, F32)]), fns))
            dx = acrobot_drift(x, u, params)
            da = np.array([foriloop(tr, np.concatenate([y, a, [u]]).astype(F32), fns) for tr in state_eq], dtype=F32)
            return np.concatenate([dx, da]).astype(F32)

        y = np.concatenate([x0, np.zeros(S, F32)]).astype(F32)
        t1 = F32(ts[-1])
        tprev, tnext = F32(ts[0]), min(F32(ts[0] + F32(dt0)), t1)
        saved = np.full((len(ts), 4 + S), np.inf, dtype=F32)
        k = 0
        while tprev < t1:
            y1 = (y + (tnext - tprev) * drift(tprev, y)).astype(F32)
            while k < len(ts) and ts[k] <= tnext:
                c = F32(0) if tprev == tnext else F32((ts[k] - tprev) / (tnext - tprev))
                saved[k] = y + c * (y1 - y)
                k += 1
            bad = (abs(y1[2]) > F32(8 * np.pi)) or (abs(y1[3]) > F32(18 * np.pi)) or not np.isfinite(y1).all()
            y, tprev = y1, min(tnext, t1)
            nn = F32(tnext + F32(dt0))
            tnext = t1 if nn > F32(t1 - F32(1e-6)) else nn
            if bad:
                break
        xs, acts = saved[:, :4], saved[:, 4:]
        us = np.array([foriloop(readout[0], np.concatenate([acrobot_obs(xs[i], noise_at(ts[i])), acts[i],
                                                            np.zeros(1, F32)]).astype(F32), fns)
                       for i in range(len(ts))], dtype=F32)
        return xs, acts, us, acrobot_fitness(xs, us, ts)
