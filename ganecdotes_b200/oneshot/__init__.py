"""Drop-in for the two functions of the reference's `lib/oneshot/image_augmentor.py` that the clustering
path calls (ref :8-56 and :59-104), on the sm_100a generator (`ganecdotes_b200.stylegan2.model.Generator`).

The training engine (`hfc_with_swav/engine.py`) does not go through these: it batches the mapping network over
both views and skips the ten `model.style(randn)` passes the reference multiplies by sigma = 0.  These functions
keep the reference's call signatures and return layout for external callers (support-set generation,
`create_hidden_features_from_perturbed_vectors`), and consume the CPU random stream exactly as the
reference does on a CPU tensor (one `randn_like` per W+ row).
"""
import torch


@torch.no_grad()
def create_perturbed_vectors_from_latents(input_latents, model, n_samples=10, n_layers=6, perturb_std=[0.25] * 6):
    """ref image_augmentor.py:8-56.  `input_latents` [1, n_latent, D] (W+); returns a list of 2*n_layers tensors
    [n_samples, D]: row n of W+ blended with `model.style(randn)` by perturb_std[n].

    Rows with perturb_std[n] == 0 come back as plain copies; their random draws are still consumed (on the
    generator of `input_latents`' device, like `torch.randn_like` in the reference), but the mapping-network
    pass the reference multiplies by zero is skipped."""
    out = []
    for n in range(2 * n_layers):
        cur = input_latents[0, n, :].clone()
        rows = cur.repeat(n_samples, 1)
        z = torch.randn_like(rows)                    # consumed whether or not the row is perturbed
        sg = float(perturb_std[n])
        if sg == 0.0:
            out.append(rows)
            continue
        noises = model.style(z.to(next(model.parameters()).device).float().contiguous()).to(rows.device)
        out.append((1 - sg) * rows + sg * noises)
    return out


@torch.no_grad()
def create_images_and_features_from_perturbed_latents(perturbations, model, model_args, layer_no=None,
                                                      return_image=True, return_feat=True, skip_const=False):
    """ref image_augmentor.py:59-104.  `perturbations` [B, n_latent, D] (W+), truncated once more with
    model_args['truncation'] / ['mean_latent'] (the reference's double truncation), fixed
    noise buffers.  Features are regrouped 13 -> 7 (or 6 with skip_const): [F0] + [cat(F_{2n+1}, F_{2n+2})]."""
    imgs, feats = model([perturbations], truncation=model_args['truncation'],
                        truncation_latent=model_args['mean_latent'], input_is_latent=True, randomize_noise=False)
    n_layers = len(feats) // 2
    grouped = [torch.cat([feats[2 * n + 1], feats[2 * n + 2]], 1) for n in range(n_layers)]
    if not skip_const:
        grouped = [feats[0]] + grouped
    picked = grouped if layer_no is None else grouped[layer_no]
    if return_feat and return_image:
        return imgs, picked
    if return_image:
        return imgs
    if return_feat:
        return picked
    return None


__all__ = ["create_perturbed_vectors_from_latents", "create_images_and_features_from_perturbed_latents"]
