"""Drop-in command line for the generation path: `python -m sample.generate_{uncond,cat,sketch,image,text}` with the
reference's flags (utils/parser_util.py:40-170: --model_path --output_dir --cond_mode --ae_dir --num_samples --resolution
--guidance_param --device --batch_size --seed --category --sketch_path --image_path --mask_path --prompt --watertight
--noise_schedule --sigma_small --cond_mask_prob --dataset ...) and checkpoint layouts.

What runs here: reverse diffusion -> UDF lattice -> MeshUDF marching cubes -> UDF face filter (the hot path), then the
reference's mesh clean-up (meshudf.py:379-434, surfd_b200/meshclean.py) and output stage (Laplacian smoothing, removal of
components under 2500 faces, .obj; generate_uncond.py:113-122, surfd_b200/output.py) on the device -- both restated from the
documented behaviour of trimesh / pymeshlab, which are absent here (parity unpinned).  The CLIP ViT-B/32 conditioning step of the text / image /
sketch scripts runs on the device through surfd_b200/clip_encoder.py (--clip_path: the weights are not shipped; --clip_vocab: the
BPE merges file for text), once per generation; --context_path takes pre-computed 512-d embeddings instead.
`--watertight` (generate_image.py / generate_text.py:132-158: udf lattice -> `mcubes.marching_cubes(udf, 0.01)` -> components
under 5000 faces removed) runs on the device through surfd_b200/watertight.py for the two scripts that have the flag
(parity unpinned: PyMCubes absent); the other three scripts parse and ignore it, like the reference.
File names follow the scripts: {k}.obj, <category name>/{k}.obj, sketch_<sketch file stem>.obj, <image file stem>.obj,
<prompt with dashes>_{k}.obj (the sketch / image scripts produce ONE shape; with --num_samples > 1 a _{k} suffix is added).
Extra flags: --context_path, --clip_path, --clip_vocab, --dense_grid (use_fast_grid_filler=False), --precision {fp32,tf32}, --raw_mesh (stop at the
meshudf.py:379 boundary).
Multi-GPU: launch with torchrun; samples are sharded contiguously over ranks; rank 0 alone reads the two checkpoints and
packs them, the flat blobs reach the other ranks through ONE NCCL broadcast each (surfd_b200.dist.broadcast_packed).
"""
import argparse
import os
import time

import torch


def generate_args(argv=None):
    p = argparse.ArgumentParser()
    g = p.add_argument_group("base")
    g.add_argument("--num_actions", default=9, type=int, help="num_classes.")
    g.add_argument("--cuda", default=True, type=bool, help="Use cuda device, otherwise use CPU.")
    g.add_argument("--device", default=0, type=int, help="Device id to use.")
    g.add_argument("--seed", default=10, type=int, help="For fixing random seed.")
    g.add_argument("--batch_size", default=64, type=int, help="Batch size during training.")
    g.add_argument("--distributed", default=False, type=bool, help="Use ddp to train model")
    g = p.add_argument_group("sampling")
    g.add_argument("--model_path", required=True, type=str, help="Path to model####.pt file to be sampled.")
    g.add_argument("--output_dir", default="", type=str, help="Path to results dir (auto created by the script).")
    g.add_argument("--num_samples", default=1, type=int, help="Maximal number of prompts to sample.")
    g.add_argument("--guidance_param", default=1.0, type=float, help="For classifier-free sampling - the s parameter.")
    g.add_argument("--if_clip", action="store_true")
    g.add_argument("--clip_value", default=0.1, type=float, help="max_clipping value (0-max).")
    g = p.add_argument_group("generate")
    g.add_argument("--grid_size", default=128, type=int, help="grid size.")
    g.add_argument("--category", default=0, type=int, help="Condition category.")
    g.add_argument("--sketch_path", default=None, type=str, help="Path to the condition sketch image.")
    g.add_argument("--image_path", default=None, type=str, help="Path to the condition image.")
    g.add_argument("--mask_path", default=None, type=str, help="Path to the condition mask.")
    g.add_argument("--prompt", default=None, type=str, help="text prompt for generation.")
    g.add_argument("--watertight", action="store_true", help="mesh attributes.")
    g.add_argument("--resolution", default=512, type=int, help="mesh resolution.")
    g.add_argument("--ae_dir", default=None, type=str, help="Path to ae")
    g = p.add_argument_group("dataset")
    g.add_argument("--dataset", default="deepfashion3d", choices=["deepfashion3d", "text2shape", "pix3d", "kcars"], type=str)
    g.add_argument("--data_dir", default="", type=str)
    g = p.add_argument_group("model")
    g.add_argument("--arch", default="OpenUNet", choices=["OpenUNet"], type=str)
    g.add_argument("--cond_mask_prob", default=0, type=float)
    g.add_argument("--unconstrained", action="store_true")
    g.add_argument("--cond_mode", choices=["no_cond", "text", "sketch", "category", "img"], type=str, required=True, help="condition type")
    g = p.add_argument_group("diffusion")
    g.add_argument("--noise_schedule", default="cosine", choices=["linear", "cosine"], type=str)
    g.add_argument("--diffusion_steps", default=1000, type=int, help="parsed and ignored like the reference (steps = 1000, SURVEY F6)")
    g.add_argument("--sigma_small", default=True, type=bool)
    g = p.add_argument_group("surfd_b200 extensions")
    g.add_argument("--context_path", default=None, type=str, help="torch file with pre-computed [B,512] CLIP embeddings")
    g.add_argument("--clip_path", default=None, type=str, help="OpenAI CLIP ViT-B/32 checkpoint (ViT-B-32.pt); default: $SURFD_CLIP_PATH, ~/.cache/clip/ViT-B-32.pt")
    g.add_argument("--clip_vocab", default=None, type=str, help="bpe_simple_vocab_16e6.txt.gz of a CLIP install (text mode)")
    g.add_argument("--dense_grid", action="store_true", help="use_fast_grid_filler=False (dense lattice)")
    g.add_argument("--precision", default="fp32", choices=["fp32", "tf32"], help="decoder GEMM precision")
    g.add_argument("--raw_mesh", action="store_true", help="write the mesh at the meshudf.py:379 boundary (no clean-up / smoothing)")
    args = p.parse_args(argv)
    if args.cond_mask_prob == 0:          # parse_and_load_from_model (utils/parser_util.py:18-19)
        args.guidance_param = 1
    return args


def write_obj(path, verts, faces):
    """minimal Wavefront writer for --raw_mesh (`v %.6f %.6f %.6f`, `f a b c`); text conversion in the C library"""
    from .output import _native_write
    _native_write(path, "# surfd_b200\n", verts.detach().to(torch.float64).cpu().numpy(), 0, "", faces.detach().cpu().numpy(), "")


CAT2NAME = {0: "long_sleeve_upper", 1: "short_sleeve_upper", 2: "no_sleeve_upper", 3: "long_sleeve_dress", 4: "short_sleeve_dress",
            5: "no_sleeve_dress", 6: "long_pants", 7: "short_pants", 8: "dress"}      # generate_cat.py:21-29


def mesh_path_for(args, kind, k, B):
    """output file of sample k, named like the reference's scripts (generate_uncond.py:114, generate_cat.py:121,
    generate_sketch.py:124,145, generate_image.py:92-94,147, generate_text.py:130)"""
    many = f"_{k}" if B > 1 else ""
    if kind == "cat":
        return os.path.join(args.output_dir, CAT2NAME.get(args.category, str(args.category)), f"{k}.obj")
    if kind == "sketch" and args.sketch_path:
        return os.path.join(args.output_dir, f"sketch_{args.sketch_path.split('/')[-1][:-4]}{many}.obj")
    if kind == "image" and args.image_path:
        return os.path.join(args.output_dir, f"{args.image_path.split('/')[-1].split('.')[0]}{many}.obj")
    if kind == "text" and args.prompt:
        return os.path.join(args.output_dir, args.prompt.replace(" ", "-").replace(".", "")[:100] + f"_{k}.obj")
    return os.path.join(args.output_dir, f"{k}.obj")


def _clip_checkpoint(args):
    for c in (args.clip_path, os.environ.get("SURFD_CLIP_PATH"), os.path.expanduser("~/.cache/clip/ViT-B-32.pt")):
        if c and os.path.exists(c):
            return c
    return None


def _context(args, kind, B, device):
    """[B, 512] conditioning: pre-computed embeddings (--context_path), or the scripts' own CLIP ViT-B/32 step on the device
    (surfd_b200/clip_encoder.py: generate_text.py:97 + mdm.py:86-97, generate_image.py:92-115, generate_sketch.py:74-82) -- once per
    generation, not once per denoiser call"""
    if args.context_path:
        ctx = torch.load(args.context_path, map_location="cpu")
        ctx = torch.as_tensor(ctx, dtype=torch.float32).reshape(-1, 512)
        if ctx.shape[0] == 1:
            ctx = ctx.repeat(B, 1)
        if ctx.shape[0] < B:
            raise SystemExit(f"--context_path holds {ctx.shape[0]} embeddings but {B} samples were requested")
        return ctx[:B].contiguous()
    ckpt = _clip_checkpoint(args)
    if ckpt is None:
        raise SystemExit(f"cond_mode={kind}: no CLIP ViT-B/32 checkpoint (weights are not shipped): pass --clip_path / set SURFD_CLIP_PATH, "
                         "or pass --context_path with pre-computed [B,512] embeddings")
    from . import clip_encoder as CE
    enc = CE.ClipEncoder.from_file(ckpt, device)
    if kind == "text":
        if not args.prompt:
            raise SystemExit("--prompt is required for generate_text")
        tok = CE.Tokenizer(CE.find_vocab(args.clip_vocab or ckpt))
        return enc.encode_text(tok.tokenize([args.prompt] * B, truncate=True)).float()
    if kind == "image":
        if not (args.image_path and args.mask_path):
            raise SystemExit("--image_path and --mask_path are required for generate_image")
        return enc.encode_image(CE.image_condition(args.image_path, args.mask_path)).float().repeat(B, 1)
    if not args.sketch_path:
        raise SystemExit("--sketch_path is required for generate_sketch")
    return enc.encode_image(CE.sketch_condition(args.sketch_path)).float().repeat(B, 1)


def main(kind, argv=None):
    """kind in {'uncond','cat','sketch','image','text'}; mirrors sample/generate_*.py main()."""
    import torch.distributed as dist
    from .pipeline import SurfDPipeline
    args = generate_args(argv)
    out_path = args.output_dir
    os.makedirs(out_path, exist_ok=True)
    assert args.num_samples <= args.batch_size, \
        f"Please either increase batch_size({args.batch_size}) or reduce num_samples({args.num_samples})"
    args.batch_size = args.num_samples
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(args.device)))
    if not torch.cuda.is_available():
        raise SystemExit("surfd_b200 has no CPU path: a CUDA device is required")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    watertight = args.watertight and kind in ("image", "text")      # the other three scripts never read the flag
    if not args.sigma_small:
        raise NotImplementedError("--sigma_small False (FIXED_LARGE variance) is never used by the reference's checkpoints")
    latent = 64 if kind in ("image", "text") else 32      # generate_image.py:76, generate_text.py:80 vs generate_uncond.py:55
    cond_mode = args.cond_mode
    from . import _lib, unet as U
    from .decoder import pack_decoder
    from .dist import broadcast_packed, shard_range
    a = U.arch(latent, cond_mode, args.num_actions)
    if rank == 0:
        print("Creating model and diffusion...")
        print(f"Loading checkpoints from [{args.model_path}]...")
        state = torch.load(args.model_path, map_location="cpu")
        blob_u, prog, _ = U.pack_unet(state, latent, cond_mode, args.num_actions)
        ckpt = torch.load(args.ae_dir, map_location="cpu")
        print(f"Load AutoEncoder From: {args.ae_dir}")
        blob_d = pack_decoder(ckpt["decoder"], latent)
        blob_u, prog, blob_d = blob_u.to(dev), prog.to(dev), blob_d.to(dev)
    else:
        blob_u = torch.empty(a.n_floats, dtype=torch.float32, device=dev)
        prog = torch.empty(16 + len(a.buffers) + len(a.prog) * U.REC, dtype=torch.int64, device=dev)
        blob_d = torch.empty(_lib.load().surfd_dec_packed_floats(latent), dtype=torch.float32, device=dev)
    broadcast_packed([blob_u, prog, blob_d], src=0)
    B = args.batch_size
    lo, hi = shard_range(B, world, rank)
    per = max(1, hi - lo)
    pipe = SurfDPipeline(None, None, latent, cond_mode, device=dev, max_batch=per, num_actions=args.num_actions,
                         mc_parallel=min(8, per), packed_unet=(blob_u, prog.cpu()), packed_decoder=blob_d)
    if args.precision == "tf32":
        pipe.decoder.set_precision(1)
    # noise: CPU generator, full-batch order, sliced per rank (identical for any GPU count)
    gen = torch.Generator().manual_seed(args.seed)
    noise = torch.randn(1001, B, latent, generator=gen)[:, lo:hi].contiguous()
    ctx = lab = None
    if kind in ("sketch", "image", "text"):
        from .dist import conditioning_from_rank0
        ctx = conditioning_from_rank0(lambda: _context(args, kind, B, dev), B, 512, dev, rank, world)[lo:hi]
    if kind == "cat":
        if not 0 <= args.category < args.num_actions:
            raise IndexError(f"--category {args.category} outside [0, {args.num_actions})")
        lab = torch.full((B,), args.category, dtype=torch.int64)[lo:hi]
    t0 = time.time()
    if hi > lo:
        from .meshclean import clean_mesh
        from .output import finish_and_save, write_obj_meshlab
        if watertight:
            lat = pipe.sample_latents(noise.to(dev), ctx, lab, float(args.guidance_param), 1000, args.noise_schedule)
            meshes = pipe.watertight(lat, args.resolution, iso=0.01, mincomponentsize=5000)
        else:
            lat, meshes, stats = pipe.generate(noise.to(dev), args.resolution, ctx, lab, guidance=float(args.guidance_param),
                                               n_steps=1000, use_fast_grid_filler=not args.dense_grid, noise_schedule=args.noise_schedule)
        torch.cuda.synchronize()
        for k, (v, f) in enumerate(meshes):
            mesh_path = mesh_path_for(args, kind, lo + k, B)
            os.makedirs(os.path.dirname(mesh_path) or ".", exist_ok=True)
            if watertight:
                write_obj_meshlab(mesh_path, v, f)
            elif args.raw_mesh:
                write_obj(mesh_path, v, f)
            else:
                v, f = clean_mesh(v, f, smooth_borders=True)
                finish_and_save(v, f, mesh_path, mincomponentsize=2500)
        print(f"rank {rank}: {hi - lo} shapes in {time.time() - t0:.2f}s; saved results to {mesh_path}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
