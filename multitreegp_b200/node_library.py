"""Node library: node id -> opcode table, rebuilt exactly like the reference assigns ids.

Reference: GeneticProgramming.__init__, genetic_programming.py:132-199 --
  id 0 = empty row, id 1 = coefficient leaf, ids 2.. = operators in ``operator_list`` order
  de-duplicated by display string (:143-162), then variables in ``variable_list`` order across
  layers, de-duplicated, the j-th NEW variable reading ``data[j]`` (:171-180); ``slots`` is the
  arity per id (:195); ``variable_array[tree, j]`` marks which variables a tree may sample (:183-193).

In the reference the node functions are arbitrary Python lambdas; the CUDA interpreter has a
closed instruction set, so operators are mapped by their display string (``"+"``, ``"sin"`` ...)
and anything else raises at construction time -- there is no CPU fallback.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from . import abi


class NodeLibrary:
    def __init__(self, operator_list: Sequence[tuple], variable_list: Sequence[Sequence[str]],
                 layer_sizes: Sequence[int], opcode_map: dict | None = None) -> None:
        layer_sizes = [int(v) for v in np.asarray(layer_sizes).reshape(-1)]
        assert len(operator_list) > 0, "No operators were given"
        assert len(layer_sizes) == len(variable_list), "There is not a set of expressions for every type of layer"
        known = dict(abi.OPERATOR_STRINGS)
        if opcode_map:
            known.update(opcode_map)

        self.string_to_node: dict[str, int] = {}
        self.node_to_string: dict[int, str] = {}
        opcodes = [abi.OP_EMPTY, abi.OP_COEF]
        args = [0, 0]
        n_operands = [0, 0]
        probabilities = []
        index = 2
        for operator_tuple in operator_list:
            string = operator_tuple[0]
            # tuples are (string, callable, arity[, probability]); the callable is ignored here
            if len(operator_tuple) >= 3 and callable(operator_tuple[1]):
                arity = operator_tuple[2]
                probability = operator_tuple[3] if len(operator_tuple) == 4 else 1.0
            else:  # (string, arity[, probability]) shorthand for callers without JAX lambdas
                arity = operator_tuple[1]
                probability = operator_tuple[2] if len(operator_tuple) == 3 else 1.0
            if string in self.string_to_node:
                continue
            if string not in known:
                raise ValueError(
                    f"operator {string!r} is not in the sm_100a interpreter's instruction set "
                    f"({sorted(known)}); pass opcode_map={{name: (opcode, arity)}} to alias it")
            opcode, op_arity = known[string]
            if int(arity) != op_arity:
                raise ValueError(f"operator {string!r} declared with arity {arity}, instruction has arity {op_arity}")
            self.string_to_node[string] = index
            self.node_to_string[index] = string
            opcodes.append(opcode)
            args.append(0)
            n_operands.append(op_arity)
            probabilities.append(float(probability))
            index += 1
        self.operator_indices = np.arange(2, index)
        self.operator_probabilities = np.asarray(probabilities, dtype=np.float64)
        var_start_index = index

        data_index = 0
        for var_list in variable_list:
            assert len(var_list) > 0, "An empty set of variables was given"
            for var in var_list:
                if var not in self.string_to_node:
                    self.string_to_node[var] = index
                    self.node_to_string[index] = var
                    opcodes.append(abi.OP_VAR)
                    args.append(data_index)
                    n_operands.append(0)
                    index += 1
                    data_index += 1
        self.variable_indices = np.arange(var_start_index, index)
        self.n_data = data_index
        self.layer_sizes = layer_sizes
        self.num_trees = int(sum(layer_sizes))

        variable_array = np.zeros((self.num_trees, data_index))
        counter = 0
        for layer_i, var_list in enumerate(variable_list):
            p = np.zeros(data_index)
            for var in var_list:
                p[self.string_to_node[var] - var_start_index] = 1
            for _ in range(layer_sizes[layer_i]):
                variable_array[counter] = p
                counter += 1
        self.variable_array = variable_array
        self.slots = np.asarray(n_operands)
        self.opcodes = np.asarray(opcodes, dtype=np.uint8)
        self.args = np.asarray(args, dtype=np.int32)
        if index > abi.MTGP_MAX_NODE_IDS:
            raise ValueError(f"{index} node ids exceed MTGP_MAX_NODE_IDS={abi.MTGP_MAX_NODE_IDS}")

    @property
    def n_ids(self) -> int:
        return len(self.opcodes)

    def table(self) -> abi.NodeTable:
        t = abi.NodeTable()
        t.n_ids = self.n_ids
        for i, (o, a) in enumerate(zip(self.opcodes, self.args)):
            t.opcode[i] = int(o)
            t.arg[i] = int(a)
        return t
