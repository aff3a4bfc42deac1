"""Engine: the lockstep batched MCTS / self-play engine handle (rvs_engine_* in rvs_b200.h)."""
import ctypes as C

import numpy as np

from . import _lib as L


class Engine:
    """One engine per device.  Not thread-safe (one host thread per handle)."""

    def __init__(self, n_games, max_sims, max_wave=64, evaluator=L.EVAL_E0, c_puct=1.0, rules=L.RULES_REF,
                 seed=0, device=None, nodes_per_game=0, net_blocks=0, net_filters=0, sample_capacity=0):
        # device=None: the process's current CUDA device (torch.cuda.current_device()), so that a torchrun rank
        # that called torch.cuda.set_device(local_rank) gets an engine on ITS GPU; the library restores the
        # caller's current device after every call
        if not isinstance(device, (int, np.integer)):  # None, 'cuda', 'cuda:1', torch.device
            d = "" if device is None else str(device)
            device = int(d.split(":", 1)[1]) if ":" in d else L.current_device()
        device = int(device)
        cfg = L.EngineConfig(C.sizeof(L.EngineConfig), device, n_games, max_sims, max_wave, rules, evaluator,
                             c_puct, seed, nodes_per_game, net_blocks, net_filters, sample_capacity)
        self._h = C.c_void_p()
        L.check(L.lib().rvs_engine_create(C.byref(cfg), C.byref(self._h)))
        self.n_games, self.max_sims, self.max_wave = n_games, max_sims, max_wave
        self.evaluator, self.rules, self.device = evaluator, rules, device
        self.sample_capacity = sample_capacity if sample_capacity > 0 else 64 * n_games
        self.cur_k = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib().rvs_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _s(self, stream):
        # torch's current stream ON THE ENGINE'S DEVICE (a stream of another device would be an invalid handle)
        return stream if stream is not None else L.current_stream(self.device)

    def _check_device(self, *tensors):
        for t in tensors:
            if hasattr(t, "is_cuda") and t.is_cuda and t.device.index != self.device:
                raise ValueError(f"tensor on {t.device} passed to an engine on cuda:{self.device}")

    def reset(self, stream=None):
        L.check(L.lib().rvs_engine_reset(self._h, self._s(stream)))

    def set_positions(self, black, white, side, stream=None):
        mem = L.mem_of(black, white, side)
        self._check_device(black, white, side)
        L.check(L.lib().rvs_engine_set_positions(self._h, L.ptr(black, "uint64")[0], L.ptr(white, "uint64")[0],
                                                 L.ptr(side, "uint8")[0], len(black), mem, self._s(stream)))

    def get_positions(self, n=None, stream=None):
        n = self.n_games if n is None else n
        b = np.empty(n, dtype=np.uint64)
        w = np.empty(n, dtype=np.uint64)
        s = np.empty(n, dtype=np.uint8)
        f = np.empty(n, dtype=np.uint8)
        L.check(L.lib().rvs_engine_get_positions(self._h, b.ctypes.data, w.ctypes.data, s.ctypes.data,
                                                 f.ctypes.data, n, L.MEM_HOST, self._s(stream)))
        return b, w, s, f

    def search(self, num_sims, wave, stream=None):
        L.check(L.lib().rvs_engine_search(self._h, num_sims, wave, self._s(stream)))

    # ---- external-evaluator path (MCTS._traverse / _process_batch split at model.predict) ----
    def begin_search(self, stream=None):
        L.check(L.lib().rvs_engine_begin_search(self._h, self._s(stream)))

    def select(self, k, stream=None):
        L.check(L.lib().rvs_engine_select(self._h, k, self._s(stream)))
        self.cur_k = k

    def leaf_planes(self, device=None, stream=None):
        """([n_games*k,3,8,8] f32, [n_games*k] uint8 valid) as numpy (device=None) or torch CUDA tensors"""
        n = self.n_games * self.cur_k
        if device is None:
            planes = np.empty((n, 3, 8, 8), dtype=np.float32)
            valid = np.empty(n, dtype=np.uint8)
            mem = L.MEM_HOST
        else:
            import torch
            planes = torch.empty((n, 3, 8, 8), dtype=torch.float32, device=device)
            valid = torch.empty(n, dtype=torch.uint8, device=device)
            mem = L.MEM_DEVICE
        L.check(L.lib().rvs_engine_leaf_planes(self._h, L.ptr(planes)[0], L.ptr(valid)[0], mem, self._s(stream)))
        return planes, valid

    def process(self, probs, values, stream=None):
        mem = L.mem_of(probs, values)
        self._check_device(probs, values)
        L.check(L.lib().rvs_engine_process(self._h, L.ptr(probs, "float32")[0], L.ptr(values, "float32")[0], mem, self._s(stream)))
        self.cur_k = 0

    def root_visits(self, n=None, device=None, stream=None):
        n = self.n_games if n is None else n
        if device is None:
            out = np.empty((n, 65), dtype=np.int32)
            mem = L.MEM_HOST
        else:
            import torch
            out = torch.empty((n, 65), dtype=torch.int32, device=device)
            mem = L.MEM_DEVICE
        L.check(L.lib().rvs_engine_root_visits(self._h, L.ptr(out)[0], n, mem, self._s(stream)))
        return out

    def play(self, temperature=1.0, recycle=True, want_moves=False, stream=None):
        mv = np.empty(self.n_games, dtype=np.uint8) if want_moves else None
        L.check(L.lib().rvs_engine_play(self._h, float(temperature), 1 if recycle else 0,
                                        None if mv is None else mv.ctypes.data, L.MEM_HOST, self._s(stream)))
        return mv

    def selfplay(self, num_sims, plies, temperature=1.0, recycle=True, stream=None):
        """persistent self-play launch: `plies` game-plies in total over all slots (wave 1)"""
        L.check(L.lib().rvs_engine_selfplay(self._h, num_sims, float(temperature), int(plies), 1 if recycle else 0,
                                            self._s(stream)))

    def drain_samples(self, capacity=None, device=None, stream=None):
        """completed-game samples (states [n,3,8,8] f32, pi [n,65] f32, z [n] f32)"""
        cap = capacity if capacity is not None else self.sample_capacity
        cnt = C.c_int64(0)
        if device is None:
            st = np.empty((cap, 3, 8, 8), dtype=np.float32)
            pi = np.empty((cap, 65), dtype=np.float32)
            z = np.empty(cap, dtype=np.float32)
            mem = L.MEM_HOST
        else:
            import torch
            st = torch.empty((cap, 3, 8, 8), dtype=torch.float32, device=device)
            pi = torch.empty((cap, 65), dtype=torch.float32, device=device)
            z = torch.empty(cap, dtype=torch.float32, device=device)
            mem = L.MEM_DEVICE
        L.check(L.lib().rvs_engine_drain_samples(self._h, L.ptr(st)[0], L.ptr(pi)[0], L.ptr(z)[0], cap,
                                                 C.byref(cnt), mem, self._s(stream)))
        n = cnt.value
        return st[:n], pi[:n], z[:n]

    def drain_packed(self, capacity=None, device=None, stream=None):
        """completed-game samples as the engine keeps them (replay.PackedSamples: black, white, side,
        z int8, pi f32[65]; 277 B per sample) -- numpy on the host, or torch tensors on `device`"""
        from .replay import PackedSamples
        cap = capacity if capacity is not None else self.sample_capacity
        cnt = C.c_int64(0)
        if device is None:
            bl = np.empty(cap, dtype=np.uint64); wh = np.empty(cap, dtype=np.uint64)
            sd = np.empty(cap, dtype=np.uint8); z = np.empty(cap, dtype=np.int8)
            pi = np.empty((cap, 65), dtype=np.float32)
            mem = L.MEM_HOST
        else:
            import torch
            bl = torch.empty(cap, dtype=torch.int64, device=device); wh = torch.empty(cap, dtype=torch.int64, device=device)
            sd = torch.empty(cap, dtype=torch.uint8, device=device); z = torch.empty(cap, dtype=torch.int8, device=device)
            pi = torch.empty((cap, 65), dtype=torch.float32, device=device)
            mem = L.MEM_DEVICE
        L.check(L.lib().rvs_engine_drain_packed(self._h, L.ptr(bl)[0], L.ptr(wh)[0], L.ptr(sd)[0], L.ptr(z)[0], L.ptr(pi)[0], cap,
                                                C.byref(cnt), mem, self._s(stream)))
        n = cnt.value
        return PackedSamples(bl[:n], wh[:n], sd[:n], z[:n], pi[:n])

    def set_root_noise(self, alpha, epsilon):
        """Dirichlet(alpha) noise mixed into the root priors with weight epsilon by every following
        search (engine feature: the reference configures it, src/config.py:25-26, but never applies
        it); epsilon = 0 switches it off"""
        L.check(L.lib().rvs_engine_set_root_noise(self._h, float(alpha), float(epsilon)))

    def set_option(self, option, value):
        """rvs_engine_set_option (L.OPT_*): search mode REF/FAST, game limit, CUDA-graph replay, tower grid cap ..."""
        L.check(L.lib().rvs_engine_set_option(self._h, int(option), int(value)))

    def set_search_mode(self, mode):
        """L.MODE_REF: the reference's wave semantics (graded); L.MODE_FAST: effective virtual-loss leaf batching"""
        self.set_option(L.OPT_SEARCH_MODE, mode)

    def drain_packed_async(self, capacity, device, count_out=None, stream=None):
        """rvs_engine_drain_packed_async: no host synchronisation.  Returns (PackedSamples of `capacity` rows on
        `device`, count tensor); only the first count[0] rows are samples -- read the count after synchronising
        `stream` (count_out: a pinned host or device int64[1] tensor; default a fresh device tensor)"""
        import torch
        from .replay import PackedSamples
        cap = int(capacity)
        bl = torch.empty(cap, dtype=torch.int64, device=device); wh = torch.empty(cap, dtype=torch.int64, device=device)
        sd = torch.empty(cap, dtype=torch.uint8, device=device); z = torch.empty(cap, dtype=torch.int8, device=device)
        pi = torch.empty((cap, 65), dtype=torch.float32, device=device)
        cnt = count_out if count_out is not None else torch.zeros(1, dtype=torch.int64, device=device)
        self._check_device(bl)
        L.check(L.lib().rvs_engine_drain_packed_async(self._h, bl.data_ptr(), wh.data_ptr(), sd.data_ptr(), z.data_ptr(),
                                                      pi.data_ptr(), cap, cnt.data_ptr(), self._s(stream)))
        return PackedSamples(bl, wh, sd, z, pi), cnt

    def set_lanes_per_game(self, lanes):
        """wave-1 kernels: lanes of a warp per game (8 / 4 / 2, 0 = automatic); never changes results"""
        L.check(L.lib().rvs_engine_set_lanes_per_game(self._h, int(lanes)))

    def stats(self, stream=None):
        st = L.EngineStats()
        L.check(L.lib().rvs_engine_stats_get(self._h, C.byref(st), self._s(stream)))
        return {k: getattr(st, k) for k, _ in L.EngineStats._fields_}

    def load_weights(self, flat, stream=None):
        """flat f32 state_dict (network.pack_state_dict); BN folding + bf16 packing happen on device"""
        L.check(L.lib().rvs_engine_load_weights(self._h, L.ptr(flat)[0], flat.numel() if hasattr(flat, "numel") else flat.size,
                                                L.ptr(flat)[1], self._s(stream)))

    def predict(self, black, white, side, stream=None, probs=False):
        """AlphaZeroNetwork.predict on packed positions -> (logits [n,65] f32, value [n] f32); probs=True returns
        the head kernel's softmax(logits) instead (rvs_engine_predict_probs: what the built-in NN search consumes)"""
        n = len(black)
        if hasattr(black, "data_ptr"):
            import torch
            logits = torch.empty((n, 65), dtype=torch.float32, device=black.device)
            value = torch.empty(n, dtype=torch.float32, device=black.device)
        else:
            logits = np.empty((n, 65), dtype=np.float32)
            value = np.empty(n, dtype=np.float32)
        mem = L.mem_of(black, white, side, logits, value)
        self._check_device(black, white, side)
        fn = L.lib().rvs_engine_predict_probs if probs else L.lib().rvs_engine_predict
        L.check(fn(self._h, L.ptr(black, "uint64")[0], L.ptr(white, "uint64")[0], L.ptr(side, "uint8")[0], n,
                   L.ptr(logits)[0], L.ptr(value)[0], mem, self._s(stream)))
        return logits, value
