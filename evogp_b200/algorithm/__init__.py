from .selection import BaseSelection, DefaultSelection, TournamentSelection, TruncationSelection  # noqa: F401
from .crossover import BaseCrossover, DefaultCrossover  # noqa: F401
from .mutation import BaseMutation, DefaultMutation, DeleteMutation, HoistMutation, vmap_subtree  # noqa: F401
from .genetic_programming import GeneticProgramming, ParetoFront  # noqa: F401
from .fused import FusedGeneticProgramming, GraphedGeneration  # noqa: F401
