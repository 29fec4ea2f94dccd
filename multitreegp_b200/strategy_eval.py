"""The evaluation half of the reference strategy object, for callers without JAX.

Mirrors ``GeneticProgramming.evaluate_population`` (genetic_programming.py:403-433) statement by
statement -- flatten the islands (:412), evaluate (``jit_eval``, :415), add the parsimony term (:424),
record the best solution / fitness of the generation (:426-431), reshape back (:433) -- with the CUDA
evaluator in place of ``jit_eval`` and, when torch.distributed is initialised, the population sharded
over the ranks like the reference's 1-D device mesh (:255-262).  The evolutionary half (``evolve``,
genetic operators) is out of scope and stays the reference's.

Coefficient optimisation (:418-422) differentiates through the evaluator and is not available with
a CUDA evaluator; the reference's default (``coefficient_optimisation=False``) is the only mode.
"""
from __future__ import annotations

import numpy as np


class EvaluationDriver:
    def __init__(self, num_generations: int, population_size: int, fitness_function, num_populations: int = 1,
                 size_parsinomy: float = 0.0, coefficient_optimisation: bool = False) -> None:
        assert num_populations > 0, "The number of populations should be larger than 0"
        assert population_size > 0 and population_size % 2 == 0, \
            "The population_size should be larger than 0 and an even number"
        assert num_generations > 0, "The number of generations should be larger than 0"
        if coefficient_optimisation:
            raise NotImplementedError("coefficient_optimisation differentiates through the evaluator "
                                      "(genetic_programming.py:253,447); a CUDA evaluator has no VJP")
        self.num_populations = num_populations
        self.population_size = population_size
        self.fitness_function = fitness_function
        self.size_parsinomy = size_parsinomy          # (sic) the reference's spelling
        self.current_generation = 0
        self.best_fitnesses = np.zeros(num_generations, dtype=np.float32)
        self.best_solutions = None

    def evaluate_population(self, populations, data):
        """populations f32[num_populations, population_size, T, N, 4] (numpy or CUDA torch tensor)
        -> (fitness [num_populations, population_size], populations)."""
        import torch.distributed as dist
        is_np = isinstance(populations, np.ndarray)
        flat = populations.reshape((self.num_populations * self.population_size,) + tuple(populations.shape[2:]))
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from .distributed import evaluate_population_sharded
            fitness = evaluate_population_sharded(flat, data, self.fitness_function.evaluate_population,
                                                  self.size_parsinomy)
        else:
            fitness = self.fitness_function.evaluate_population(flat, data, self.size_parsinomy)
        best = int(fitness.argmin())
        best_solution = flat[best]
        if self.best_solutions is None:
            shape = (len(self.best_fitnesses),) + tuple(flat.shape[1:])
            self.best_solutions = np.zeros(shape, dtype=np.float32)
        self.best_solutions[self.current_generation] = best_solution if is_np else best_solution.cpu().numpy()
        self.best_fitnesses[self.current_generation] = float(fitness[best])
        return fitness.reshape((self.num_populations, self.population_size)), populations

    def next_generation(self) -> None:
        """The reference increments ``current_generation`` inside ``evolve`` (genetic_programming.py:496)."""
        self.current_generation += 1

    def get_statistics(self, generation: int | None = None):
        if generation is not None:
            return self.best_fitnesses[generation], self.best_solutions[generation]
        return self.best_fitnesses, self.best_solutions
