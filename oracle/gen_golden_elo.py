"""Generates tests/golden/elo.json by running the UNMODIFIED reference ELORatingSystem
(/root/reference/src/arena/arena.py:19-135) on a seeded sequence of results, plus the schedule
order of Arena.run_tournament (arena.py:328-341).  Run in the build container only:
    python oracle/gen_golden_elo.py
"""
import json
import os
import random
import sys

sys.path.insert(0, "/root/reference")
os.environ["CUDA_VISIBLE_DEVICES"] = ""
from src.arena.arena import ELORatingSystem  # noqa: E402

rng = random.Random(20261018)
players = ["alpha", "beta", "gamma", "delta", "random"]
elo = ELORatingSystem()
elo.add_player("gamma", 1650.0)
games = []
for _ in range(400):
    a, b = rng.sample(players, 2)
    s = rng.choice([1.0, 0.5, 0.0])
    rec = elo.update_ratings(a, b, s)
    games.append({"a": a, "b": b, "score_a": s, "ra": rec["rating_a_after"], "rb": rec["rating_b_after"]})
sched = []
ids = ["p0", "p1", "p2", "p3"]
for r in range(3):
    for i in range(len(ids)):
        for j in range(i + 1, len(ids)):
            p1, p2 = ids[i], ids[j]
            if (i + j + r) % 2 == 0:
                p1, p2 = p2, p1
            sched.append([r, p1, p2])
out = {"games": games, "final": elo.ratings, "games_played": elo.games_played,
       "leaderboard": [x["player_id"] for x in elo.get_leaderboard()],
       "expected": [[ra, rb, elo.get_expected_score(ra, rb)] for ra, rb in [(1500.0, 1500.0), (1500.0, 1900.0), (1723.5, 1411.25), (0.0, 3000.0)]],
       "schedule": sched}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "elo.json")
json.dump(out, open(path, "w"))
print("wrote", path, len(games))
