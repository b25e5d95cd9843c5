"""StandardPipeline — evaluate / step loop with simple logging
(reference: src/evogp/pipeline/standard.py:10-106)."""
import time

import numpy as np
import torch


class BasePipeline:
    def step(self):
        raise NotImplementedError

    def run(self):
        raise NotImplementedError


class StandardPipeline(BasePipeline):
    def __init__(self, algorithm, problem, fitness_target: float = None, generation_limit: int = 100,
                 time_limit: int = None, is_show_details: bool = True, valid_fitness_boundry: float = 1e8):
        self.algorithm = algorithm
        self.problem = problem
        self.fitness_target = fitness_target
        self.generation_limit = generation_limit
        self.time_limit = time_limit
        self.is_show_details = is_show_details
        self.valid_fitness_boundry = valid_fitness_boundry
        self.best_tree = None
        self.best_fitness = float("-inf")
        self.fitness = None

    def step(self):
        fitness = self.problem.evaluate(self.algorithm.forest)
        fitness = torch.where(torch.isnan(fitness), torch.full_like(fitness, float("-inf")), fitness)
        host = fitness.cpu()
        top = int(torch.argmax(host))
        if host[top] > self.best_fitness:
            self.best_fitness = host[top]
            self.best_tree = self.algorithm.forest[top]
        self.algorithm.step(fitness)
        return host

    def run(self):
        started = time.time()
        generation = 0
        while True:
            tic = time.time()
            self.fitness = self.step()
            if self.is_show_details:
                self.show_details(tic, generation, self.fitness)
            if self.fitness_target is not None and self.best_fitness >= self.fitness_target:
                print("Fitness target reached!")
                break
            if self.time_limit is not None and time.time() - started > self.time_limit:
                print("Time limit reached!")
                break
            generation += 1
            if generation >= self.generation_limit:
                print("Generation limit reached!")
                break
        return self.best_tree

    def show_details(self, start_time, generation_cnt, fitnesses):
        f = fitnesses.numpy()
        ok = f[(f < self.valid_fitness_boundry) & (f > -self.valid_fitness_boundry)]
        ms = (time.time() - start_time) * 1000
        if ok.size:
            stats = f"max: {ok.max():.4f}, min: {ok.min():.4f}, mean: {np.mean(ok):.4f}, std: {np.std(ok):.4f}"
        else:
            stats = "no valid fitness"
        print(f"Generation: {generation_cnt}, Cost time: {ms:.2f}ms\n", f"\tfitness: valid cnt: {ok.size}, {stats}\n")
