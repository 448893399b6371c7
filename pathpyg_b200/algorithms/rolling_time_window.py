"""``RollingTimeWindow`` (reference ``src/pathpyG/algorithms/rolling_time_window.py``): the sequence of
time-aggregated graphs of a temporal graph under a sliding window -- a caller of ``TemporalGraph.to_static_graph``,
whose weighted form merges the window's events with the library coalesce."""
from __future__ import annotations


class RollingTimeWindow:
    """Iterates ``g.to_static_graph(weighted, (t, t + window_size))`` for ``t = start_time, start_time + step_size, ...``
    while ``t <= end_time``; with ``return_window`` every item is ``(graph, (t, t + window_size))``."""

    def __init__(self, temporal_graph, window_size, step_size=1, return_window: bool = False, weighted: bool = True):
        self.g = temporal_graph
        self.window_size = window_size
        self.step_size = step_size
        self.current_time = self.g.start_time
        self.return_window = return_window
        self.weighted = weighted

    def __iter__(self) -> "RollingTimeWindow":
        return self

    def __next__(self):
        if self.current_time > self.g.end_time:
            raise StopIteration()
        window = (self.current_time, self.current_time + self.window_size)
        snapshot = self.g.to_static_graph(weighted=self.weighted, time_window=window)
        self.current_time += self.step_size
        return (snapshot, window) if self.return_window else snapshot
