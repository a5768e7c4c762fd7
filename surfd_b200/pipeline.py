"""The generation hot path end to end on one GPU: reverse diffusion -> per-shape UDF lattice -> marching cubes ->
UDF face filter, i.e. what `sample/generate_*.py` run between "x_T drawn" and the meshudf.py:379 boundary.

Shapes are independent.  The decoder work (lattice, face filter) runs on the main stream; each shape's ordered
marching-cubes replay (a single-CTA, latency-bound kernel) runs on a side stream with its own workspace, so up to
`mc_parallel` replays overlap with each other and with the next shapes' lattices.
"""
import contextlib
import time

import torch

from . import unet as U
from .decoder import UdfDecoder
from .meshudf import MarchingCubes, finish_mesh


@contextlib.contextmanager
def _nvtx(name):
    """NVTX range around a stage (sample / lattice / mc / filter): shows up in Nsight timelines, free otherwise"""
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


class SurfDPipeline:
    def __init__(self, mdm_state, ae_state, latent_dim=32, cond_mode="no_cond", device="cuda", max_batch=64, mc_parallel=8,
                 num_actions=9, packed_unet=None, packed_decoder=None):
        self.device = torch.device(device)
        self.L = latent_dim
        self.sampler = U.UNetSampler(mdm_state, latent_dim, cond_mode, num_actions, device=device, max_batch=max_batch,
                                     packed=packed_unet)
        # leave one SM per concurrent marching-cubes replay to the persistent decoder GEMM's budget and size the point chunk
        # to an EVEN number of 128-row panels per CTA (the layer-chain kernel interleaves two panels): (148 - 8) CTAs x 10 x 128 rows
        # (profiles/r2_probe_panels.log: 512^3 lattice 88.4 ms at 9 panels per CTA, 87.6 at 10, 86.9 at 16)
        n_sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        budget = max(1, n_sms - mc_parallel)
        self.decoder = UdfDecoder(ae_state, latent_dim, device=device, packed=packed_decoder, max_chunk_points=budget * 1280)
        self.decoder.set_sm_budget(budget)
        # the persistent sampler kernel (one CTA per SM, cooperative launch) leaves the same SMs free: the marching-cubes
        # replays of the previous batch keep running next to it instead of delaying its launch
        self.sampler.set_sampler(1, budget)
        self.sm_budget = budget
        # generate_many(): (sampler SMs, decoder SMs) while consecutive batches run concurrently; None = the sampler of batch
        # i+1 runs between the lattices and the face filters of batch i on one stream
        self.overlap_split = None
        self.sampler_stream = torch.cuda.Stream(device=self.device, priority=-1)
        self.mcs = [MarchingCubes(device) for _ in range(mc_parallel)]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(mc_parallel)]
        self.schedule_cache = {}

    def schedule(self, n_steps=1000, noise_schedule="cosine"):
        key = (n_steps, noise_schedule)
        if key not in self.schedule_cache:
            betas = U.cosine_betas() if noise_schedule == "cosine" else U.linear_betas()
            self.schedule_cache[key] = U.SpacedSchedule(betas, U.space_timesteps(1000, [n_steps]))
        return self.schedule_cache[key]

    def sample_latents(self, noise, context=None, labels=None, guidance=1.0, n_steps=1000, noise_schedule="cosine"):
        with _nvtx("surfd.sample"):
            return self.sampler.sample(self.schedule(n_steps, noise_schedule), noise, context, labels, guidance)

    def _launch_fields(self, wave, latents, N, use_fast_grid_filler, max_dist, marks):
        """lattice of every shape of `wave` on the current stream; each shape's marching cubes starts on its side stream
        as soon as its lattice is complete"""
        dec = self.decoder
        main = torch.cuda.current_stream(self.device)
        fields = {}
        for j, k in enumerate(wave):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
            with _nvtx("surfd.lattice"):
                dec.set_latent(latents[k])
                udf, grads, counts = dec.lattice(N, use_fast_grid_filler=use_fast_grid_filler, max_dist=max_dist)
                udf.clamp_(min=0)                                # udf[udf < 0] = 0  (meshudf.py:342)
            e1.record(main)
            marks.append((e0, e1))
            fields[k] = (udf, grads, counts)
            self.streams[j].wait_event(e1)
            with _nvtx("surfd.mc.launch"):
                self.mcs[j].launch(udf, grads, self.streams[j])
        return fields

    def _finish_fields(self, wave, latents, N, fields, meshes, stats, marks):
        """wait for each shape's replay, then mesh assembly + UDF face filter on the current stream"""
        dec = self.decoder
        main = torch.cuda.current_stream(self.device)
        for j, k in enumerate(wave):
            udf, grads, counts = fields[k]
            try:
                res = self.mcs[j].finish()
                while res is None:                               # buffers were grown: run this shape again
                    self.mcs[j].launch(udf, grads, self.streams[j])
                    res = self.mcs[j].finish()
            except Exception:
                # e.g. "No surface found": the other replays of the wave are still in flight on their streams and read
                # lattices this frame owns -- drain them before the error unwinds
                for jj in range(j + 1, len(wave)):
                    try:
                        self.mcs[jj].finish()
                    except Exception:
                        pass
                raise
            verts_raw, faces_raw = res
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
            with _nvtx("surfd.filter"):
                vertices, faces = finish_mesh(verts_raw, faces_raw, N)
                dec.set_latent(latents[k])
                keep = dec.face_filter(vertices, faces, N)
                faces_kept = faces[keep.bool()]
            e1.record(main)
            marks.append((e0, e1, "f"))
            meshes[k] = (vertices.to(torch.float32), faces_kept.to(torch.int64))
            stats[k] = dict(n_udf=counts[0], n_grad=counts[1], n_verts=int(vertices.shape[0]), n_faces_mc=int(faces.shape[0]),
                            n_faces=int(faces_kept.shape[0]), **self.mcs[j].last_stats)
            del fields[k]

    def _account(self, marks, timings):
        if timings is None:
            return
        torch.cuda.synchronize(self.device)
        for m in marks:
            key = "filter_s" if len(m) == 3 else "lattice_s"
            timings[key] = timings.get(key, 0.0) + m[0].elapsed_time(m[1]) / 1e3

    def extract(self, latents, N, use_fast_grid_filler=True, max_dist=0.1, timings=None):
        """latents [B,1,L] (device) -> list of (verts float32 [V,3], faces int64 [F,3]) + per-shape stats"""
        B = latents.shape[0]
        P = len(self.mcs)
        meshes, stats = [None] * B, [None] * B
        marks = []
        for w0 in range(0, B, P):
            wave = list(range(w0, min(B, w0 + P)))
            fields = self._launch_fields(wave, latents, N, use_fast_grid_filler, max_dist, marks)
            self._finish_fields(wave, latents, N, fields, meshes, stats, marks)
        self._account(marks, timings)
        return meshes, stats

    def watertight(self, latents, N, iso=0.01, mincomponentsize=5000):
        """the `--watertight` branch (generate_text.py:132-158): per shape, coarse-to-fine udf lattice -> classic marching
        cubes at `iso` -> components under `mincomponentsize` faces removed.  Returns [(verts [V,3] in lattice-index units,
        faces [F,3])]; the lattice stays on the device (surfd_b200/watertight.py)."""
        from .watertight import watertight_mesh
        out = []
        for k in range(latents.shape[0]):
            with _nvtx("surfd.lattice"):
                self.decoder.set_latent(latents[k])
                udf, _, _ = self.decoder.lattice(N, use_fast_grid_filler=True, grads=False)
            with _nvtx("surfd.watertight"):
                out.append(watertight_mesh(udf, iso, mincomponentsize))
            del udf
        return out

    def generate_many(self, noises, N, contexts=None, labels=None, guidance=1.0, n_steps=1000, use_fast_grid_filler=True,
                      noise_schedule="cosine", timings=None, to_host=False, io=None):
        """Several independent batches, software-pipelined on one GPU: while batch i's marching-cubes replays (one warp
        each, latency-bound) run on their side streams, the sampler of batch i+1 already runs -- on the main stream between
        batch i's lattices and face filters, or (self.overlap_split set) on its own stream and SM partition next to the
        whole extraction of batch i.  Each batch must fit one wave (B <= mc_parallel).
        noises: list of [n_steps+1, B, L] tensors, on the device or in (pinned) host memory -- host tensors are copied
        inside the loop, just before their sampler is launched.  to_host=True also copies every batch's latents/meshes
        to host memory as soon as they are complete.  `io` (dict) accumulates h2d/d2h byte counts.
        Returns a list of (latents, meshes, stats)."""
        K = len(noises)
        P = len(self.mcs)
        get = lambda seq, i: None if seq is None else seq[i]
        io = io if io is not None else {}

        def dev(t):
            if t is None or t.is_cuda:
                return t
            io["h2d"] = io.get("h2d", 0) + t.numel() * t.element_size()
            return t.to(self.device, non_blocking=True)

        def host(lat, meshes):
            outm = []
            for v, f in meshes:
                vc, fc = v.cpu(), f.cpu()
                io["d2h"] = io.get("d2h", 0) + vc.numel() * vc.element_size() + fc.numel() * fc.element_size()
                outm.append((vc, fc))
            latc = lat.cpu()
            io["d2h"] = io.get("d2h", 0) + latc.numel() * latc.element_size()
            return latc, outm

        def sample(i):
            return self.sample_latents(dev(noises[i]), dev(get(contexts, i)), dev(get(labels, i)), guidance, n_steps, noise_schedule)

        out, marks = [], []
        if any(n.shape[1] > P for n in noises):            # more shapes than replay slots: one batch after the other
            for i in range(K):
                lat = sample(i)
                meshes, stats = self.extract(lat, N, use_fast_grid_filler, timings=timings)
                out.append(host(lat, meshes) + (stats,) if to_host else (lat, meshes, stats))
            return out
        split = self.overlap_split if K > 1 else None
        if split is None:
            lat = sample(0)
            for i in range(K):
                B = lat.shape[0]
                wave = list(range(B))
                fields = self._launch_fields(wave, lat, N, use_fast_grid_filler, 0.1, marks)
                lat_next = sample(i + 1) if i + 1 < K else None
                meshes, stats = [None] * B, [None] * B
                self._finish_fields(wave, lat, N, fields, meshes, stats, marks)
                out.append(host(lat, meshes) + (stats,) if to_host else (lat, meshes, stats))
                lat = lat_next
            self._account(marks, timings)
            return out
        # Concurrent mode: the sampler of batch i+1 (one persistent kernel on `split[0]` SMs, its own stream) runs NEXT TO the
        # whole extraction of batch i (lattices, replays, face filters on the remaining SMs).  The sampler is latency-bound
        # and the extraction is interleaved with host round trips (query counts per GridFiller level, mesh sizes), so
        # each hides the other's idle time; a batch costs max(sampler, extraction) instead of their sum.
        main = torch.cuda.current_stream(self.device)
        n_samp, n_dec = split
        self.sampler.set_sampler(1, n_samp)
        self.decoder.set_sm_budget(n_dec)
        try:
            side = self.sampler_stream
            side.wait_stream(main)
            done = []
            with torch.cuda.stream(side):
                lat = sample(0)
                ev = torch.cuda.Event(); ev.record(side); done.append(ev)
            keep = [lat]
            for i in range(K):
                main.wait_event(done[i])
                lat_next = None
                if i + 1 < K:
                    with torch.cuda.stream(side):
                        lat_next = sample(i + 1)
                        ev = torch.cuda.Event(); ev.record(side); done.append(ev)
                    keep.append(lat_next)
                B = lat.shape[0]
                wave = list(range(B))
                fields = self._launch_fields(wave, lat, N, use_fast_grid_filler, 0.1, marks)
                meshes, stats = [None] * B, [None] * B
                self._finish_fields(wave, lat, N, fields, meshes, stats, marks)
                out.append(host(lat, meshes) + (stats,) if to_host else (lat, meshes, stats))
                lat = lat_next
            main.wait_stream(side)
        finally:
            self.sampler.set_sampler(1, self.sm_budget)
            self.decoder.set_sm_budget(self.sm_budget)
        self._account(marks, timings)
        return out

    def generate(self, noise, N, context=None, labels=None, guidance=1.0, n_steps=1000, use_fast_grid_filler=True,
                 noise_schedule="cosine", timings=None):
        """noise [n_steps+1, B, L] on the device.  Returns (latents [B,1,L], meshes, stats)."""
        if timings is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        lat = self.sample_latents(noise, context, labels, guidance, n_steps, noise_schedule)
        if timings is not None:
            e1.record(); torch.cuda.synchronize(self.device)
            timings["sample_s"] = timings.get("sample_s", 0.0) + e0.elapsed_time(e1) / 1e3
        meshes, stats = self.extract(lat, N, use_fast_grid_filler, timings=timings)
        return lat, meshes, stats

    def generate_host(self, noise_host, N, context_host=None, labels_host=None, **kw):
        """The call a user of the scripts makes, with HOST buffers: pinned noise/conditioning in, meshes out to host memory.
        Returns (latents_cpu, [(verts_cpu, faces_cpu)], stats, h2d_bytes, d2h_bytes)."""
        h2d = noise_host.numel() * noise_host.element_size()
        noise = noise_host.to(self.device, non_blocking=True)
        ctx = lab = None
        if context_host is not None:
            ctx = context_host.to(self.device, non_blocking=True); h2d += context_host.numel() * context_host.element_size()
        if labels_host is not None:
            lab = labels_host.to(self.device, non_blocking=True); h2d += labels_host.numel() * labels_host.element_size()
        lat, meshes, stats = self.generate(noise, N, ctx, lab, **kw)
        out, d2h = [], 0
        for v, f in meshes:
            vc, fc = v.cpu(), f.cpu()
            d2h += vc.numel() * 4 + fc.numel() * 8
            out.append((vc, fc))
        latc = lat.cpu()
        d2h += latc.numel() * 4
        return latc, out, stats, h2d, d2h
