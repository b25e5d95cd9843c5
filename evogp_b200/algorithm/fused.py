"""FusedGeneticProgramming — a generation step as ONE kernel.

Same algorithm as GeneticProgramming(DefaultSelection, DefaultCrossover, DefaultMutation) — truncation
selection with elitism, subtree crossover among the survivors, subtree mutation with freshly grown donors
(reference: algorithm/genetic_programming.py:101-120 and the three default operators) — but executed by
`torch.ops.evogp_cuda.tree_next_generation` after a single `torch.sort`: no survivor gather, no CPU mask, no
scatter, no concatenation, no host synchronisation.  Parents, positions and the mutation coin are drawn inside
the kernel (Philox4x32-10), so runs are deterministic in the torch CUDA seed but do not follow the same random
stream as the unfused operators.  Every shape is static, so a whole generation (evaluate + step) can be captured
in a CUDA graph (`capture()`)."""
import torch
from torch import Tensor

from ..tree import Forest, GenerateDescriptor

_ops = torch.ops.evogp_cuda


class FusedGeneticProgramming:
    def __init__(self, initial_forest: Forest, mutation_descriptor: GenerateDescriptor, mutation_rate: float = 0.2,
                 survival_rate: float = 0.3, elite_cnt: int = None, elite_rate: float = None):
        assert 0 <= survival_rate <= 1, "survival_rate should be in [0, 1]"
        assert elite_cnt is None or elite_rate is None, "elite_cnt and elite_rate should not be set at the same time"
        assert 0.0 <= mutation_rate <= 1.0, "mutation_rate should be in [0, 1]"
        d = mutation_descriptor
        assert (d.max_tree_len, d.input_len, d.output_len) == (initial_forest.max_tree_len, initial_forest.input_len,
                                                               initial_forest.output_len), "mutation descriptor does not match the forest"
        self.forest = initial_forest
        self.pop_size = initial_forest.pop_size
        self.descriptor = d
        self.mutation_rate = float(mutation_rate)
        self.survivor_cnt = max(1, int(self.pop_size * survival_rate))          # selection/default.py:56
        self.elite_cnt = elite_cnt if elite_cnt is not None else (int(self.pop_size * elite_rate) if elite_rate is not None else 0)
        self.generation = 0

    def step(self, fitness: Tensor, keys: Tensor = None) -> Forest:
        f = self.forest
        assert fitness.shape == (f.pop_size,), f"fitness shape should be ({f.pop_size}, ), but got {fitness.shape}"
        dev = f.batch_node_value.device
        if keys is None:   # same draw Forest.random_generate makes (forest.py:51-58)
            keys = torch.randint(low=0, high=1000000, size=(2,), dtype=torch.uint32, device=dev)
        fitness = torch.where(torch.isnan(fitness), torch.full_like(fitness, float("-inf")), fitness)
        order = torch.sort(fitness, descending=True, stable=True).indices
        d = self.descriptor
        v, t, s = _ops.tree_next_generation(f.pop_size, f.max_tree_len, f.batch_node_value.contiguous(),
                                            f.batch_node_type.contiguous(), f.batch_subtree_size.contiguous(), order,
                                            self.elite_cnt, self.survivor_cnt, self.mutation_rate, f.input_len,
                                            f.output_len, d.out_prob, d.const_prob, d.depth2leaf_probs,
                                            d.roulette_funcs, d.const_samples, keys)
        self.forest = Forest(f.input_len, f.output_len, v, t, s)
        self.generation += 1
        return self.forest


class GraphedGeneration:
    """A whole generation — evaluate the population, sort, build the next one — captured ONCE as a CUDA graph and
    replayed per generation (every shape on this path is static, nothing synchronises with the host).  The population
    lives in fixed buffers that the graph updates in place; `fitness` holds the fitness of the population that was
    evaluated by the last replay (i.e. of the PREVIOUS generation's trees).

        gen = GraphedGeneration(FusedGeneticProgramming(...), problem)
        for _ in range(100):
            gen.replay()
        best = gen.fitness.max()
    """

    def __init__(self, algorithm: FusedGeneticProgramming, problem, warmup: int = 3):
        self.algorithm = algorithm
        self.problem = problem
        f = algorithm.forest
        self.forest = Forest(f.input_len, f.output_len, f.batch_node_value.clone(), f.batch_node_type.clone(),
                             f.batch_subtree_size.clone())
        algorithm.forest = self.forest
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up off the capture stream, as torch.cuda.graph requires
            for _ in range(warmup):
                self._one_generation()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.fitness = self._one_generation()

    def _one_generation(self):
        fitness = self.problem.evaluate(self.forest)
        self.algorithm.forest = self.forest
        nxt = self.algorithm.step(fitness)
        self.forest.batch_node_value.copy_(nxt.batch_node_value)
        self.forest.batch_node_type.copy_(nxt.batch_node_type)
        self.forest.batch_subtree_size.copy_(nxt.batch_subtree_size)
        self.algorithm.forest = self.forest
        return fitness

    def replay(self):
        self.graph.replay()
        self.algorithm.generation += 1
        return self.fitness
