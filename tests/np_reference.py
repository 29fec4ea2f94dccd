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
