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
        if is_np:
            best = int(np.argmin(fitness))
            best_solution, best_fitness = flat[best], float(fitness[best])
        else:
            # device-resident generation: argmin + copy of the winner in one launch (mtgp_select);
            # the only device->host traffic is the winning candidate, once per generation
            eng = self.fitness_function._get_engine()
            _, best_f, _, best_sol = eng.select(fitness, self.num_populations, 0, flat.contiguous())
            best_solution, best_fitness = best_sol.cpu().numpy(), float(best_f.cpu()[0])
        if self.best_solutions is None:
            shape = (len(self.best_fitnesses),) + tuple(flat.shape[1:])
            self.best_solutions = np.zeros(shape, dtype=np.float32)
        self.best_solutions[self.current_generation] = best_solution
        self.best_fitnesses[self.current_generation] = best_fitness
        return fitness.reshape((self.num_populations, self.population_size)), populations

    def elite(self, fitness, elite_size: int):
        """Indices (local to each sub-population) of the ``elite_size`` fittest candidates,
        ``argsort(fitness)[:elite_size]`` of reproduction.py:82, for fitness
        [num_populations, population_size]: numpy in -> numpy out, CUDA tensor in -> CUDA tensor out."""
        if isinstance(fitness, np.ndarray):
            f = fitness.reshape(self.num_populations, self.population_size)
            return np.argsort(f, axis=1, kind="stable")[:, :elite_size].astype(np.int32)
        eng = self.fitness_function._get_engine()
        return eng.select(fitness.contiguous(), self.num_populations, elite_size)[2]

    def next_generation(self) -> None:
        """The reference increments ``current_generation`` inside ``evolve`` (genetic_programming.py:496)."""
        self.current_generation += 1

    def get_statistics(self, generation: int | None = None):
        if generation is not None:
            return self.best_fitnesses[generation], self.best_solutions[generation]
        return self.best_fitnesses, self.best_solutions
