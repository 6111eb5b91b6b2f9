"""The numbers of the reference's shipped config files for the hot path, as plain dicts
(configs/models/*.py, configs/segmentors/hfc_with_swav_*_config.py of the reference).
`swav_config(model)` returns the keyword arguments `SwAVClustering` takes (its `hfc_prep_args`),
`model_config(model)` the generator description, `seg_args(model)` the one-shot segmentor head."""
import copy
import types

_FFHQ_LAYERS = [512, 1024, 1024, 1024, 1024, 512, 256]          # layer_hf_dim, sum = 5376
_PIDRAY_LAYERS = [512, 1024, 512, 256, 128, 64, 32]             # BagGAN channel map, sum = 2528

# model name -> (generator size, is_baggan, truncation)      configs/models/{ffhq_256,lsun_*,pidray_*}.py
MODELS = {
    'ffhq-256': dict(size=256, baggan=False, truncation=0.7),
    'cat-256': dict(size=256, baggan=False, truncation=0.7),
    'afhq-256': dict(size=256, baggan=False, truncation=0.7),
    'horse-256': dict(size=256, baggan=False, truncation=0.7),
    'horse-256-rp': dict(size=256, baggan=False, truncation=0.7),
    'church-256': dict(size=256, baggan=False, truncation=0.7),
    'ffhq-256-eg': dict(size=256, baggan=False, truncation=0.7),
    'p-horse-256': dict(size=256, baggan=False, truncation=0.7),
    'p-car-512': dict(size=256, baggan=False, truncation=0.7),      # pascal_car_512.py:8 builds a 256 generator
    # lsun_car_512.py builds a 256 generator (:8,11); BASELINE.json's car-512 config asks for 512^2 features
    'car-512': dict(size=512, baggan=False, truncation=0.7),
    'pidray-256': dict(size=256, baggan=True, truncation=0.9),
}
for _n in ('pliers', 'hammer', 'powerbank', 'wrench', 'handcuffs'):     # configs/models/pidray_<tool>_256.py:8
    MODELS[f'pidray-{_n}-256'] = dict(size=256, baggan=True, truncation=0.95)

# per-method differences of hfc_with_swav_{ffhq,cat,car,horse,pidray}_config.py and the generic
# hfc_with_swav_config.py (:22 num_epochs, :52 nprototypes, :63-65 sinkhorn_args, :77-79 seg_args)
_METHOD = {
    'hfc_with_swav_ffhq': dict(nprototypes=5000, eps=0.005, source_pdf='uniform', seg='XXS', layers=_FFHQ_LAYERS,
                               num_epochs=100),
    'hfc_with_swav_cat': dict(nprototypes=5000, eps=0.003, source_pdf='image', seg='XS', layers=_FFHQ_LAYERS,
                              num_epochs=100),
    'hfc_with_swav_car': dict(nprototypes=4000, eps=0.01, source_pdf='uniform', seg='XS', layers=_FFHQ_LAYERS,
                              num_epochs=100),
    'hfc_with_swav_horse': dict(nprototypes=5000, eps=0.003, source_pdf='uniform', seg='XXS', layers=_FFHQ_LAYERS,
                                num_epochs=50),
    'hfc_with_swav_pidray': dict(nprototypes=4000, eps=0.005, source_pdf='uniform', seg='XXS', layers=_PIDRAY_LAYERS,
                                 num_epochs=100),
    'hfc_with_swav': dict(nprototypes=8000, eps=0.005, source_pdf='uniform', seg='XXS', layers=_FFHQ_LAYERS,
                          num_epochs=100),
}


def method_for(model, method='hfc_with_swav'):
    """the remap pretrain.py / evaluate.py apply to `--method` (pretrain.py:76-86)"""
    if method != 'hfc_with_swav':
        return method
    for key in ('ffhq', 'cat', 'car', 'horse'):
        if model == {'ffhq': 'ffhq-256', 'cat': 'cat-256', 'car': 'car-512', 'horse': 'horse-256'}[key]:
            return f'hfc_with_swav_{key}'
    if 'pidray' in model:
        return 'hfc_with_swav_pidray'
    return method


def model_config(model):
    m = MODELS[model]
    return types.SimpleNamespace(num_latents_for_mean=4096, truncation=m['truncation'], image_size=m['size'],
                                 latent_dim=512, is_baggan=m['baggan'],
                                 gen_args=dict(size=m['size'], style_dim=512, n_mlp=8))


def swav_config(model, method='hfc_with_swav'):
    md = copy.deepcopy(_METHOD[method_for(model, method)])
    n_hfc_layers = 6          # every shipped segmentor config (:2); the 512^2 generator is only the geometry override
    hlen = sum(md['layers'])
    return dict(
        perturb_args=dict(truncation=0.7, n_layers=n_hfc_layers, n_samples=1, layer_no=None,
                          perturb_std=[1.0] * n_hfc_layers),
        swav_args=dict(num_epochs=md['num_epochs'], num_samples=1, num_patches=5, sampling_method='random', patch_size=20000,
                       hf_interp='nearest', warmup_epochs=100, start_warmup=0.01, use_scheduler=False, base_lr=0.01,
                       final_lr=0.0001, trust_coeff=0.01, freeze_prototype_niters=313,
                       train_args=dict(lr=0.01, momentum=0.9), projn_nw='linear', temperature=0.01,
                       nprototypes=md['nprototypes'], nclasses=512, hlen=hlen, add_local_loss=False,
                       plot_test_images=False, epoch_print_freq=5, max_masks=4),
        sinkhorn_args=dict(source_pdf=md['source_pdf'], niters=10, eps=md['eps']),
        layer_hf_dim=list(md['layers']),
    )


def simclr_config(model='ffhq-256'):
    """hfc_prep_args of configs/segmentors/hfc_with_simclr_config.py (:1-48) - keyword arguments of SimCLRClustering"""
    layers = list(_FFHQ_LAYERS)
    return dict(
        perturb_args=dict(truncation=0.7, n_layers=6, n_samples=1, layer_no=None, perturb_std=[1.0] * 6),
        simclr_args=dict(num_iters=100, batch_size=20, patch_size=20000, hf_interp='nearest', trust_coeff=0.01,
                         train_args=dict(lr=0.01, momentum=0.9), temperature=1.0, nclasses=512, hlen=sum(layers),
                         epoch_print_freq=5, max_masks=4),
        layer_hf_dim=layers,
    )


KMEANS_CLUSTERS = [4, 8, 16, 32, 64]        # hfc_kmeans_config.py:6


def kmeans_config(model='ffhq-256'):
    """hfc_prep_args of configs/segmentors/hfc_kmeans_config.py (:12-40) - keyword arguments of HFCPreprocessor"""
    return dict(
        perturb_args=dict(truncation=0.7, n_layers=5, n_samples=4, perturb_std=[1.0] * 5),
        hfc_algo='hfc_kmeans',
        hfc_args=dict(kmeans_args=dict(verbose=0),
                      base_args=dict(out_dir=None, n_layers=5, clusters_per_layer=list(KMEANS_CLUSTERS), out_size=256,
                                     presaved=False)),
        hier_encode=False, hle_samples=100,
    )


def seg_args(model, method='hfc_with_swav'):
    if method == 'hfc_with_simclr':
        return dict(size='XS', in_ch=512)                       # hfc_with_simclr_config.py:55-56
    if method == 'hfc_kmeans':
        return dict(size='S', in_ch=sum(KMEANS_CLUSTERS))        # hfc_kmeans_config.py:68-69
    return dict(size=_METHOD[method_for(model, method)]['seg'], in_ch=512)


def build_generator(model, checkpoint=None, device='cuda', seed=42):
    """Generator of `model`; weights from a rosinality / BagGAN checkpoint (`g_ema` state dict,
    src/one_shot_pipeline.py:142-147) or - without one - the seeded random init the benchmarks use."""
    import torch
    from .stylegan2.model import Generator
    m = MODELS[model]
    if checkpoint is not None:
        ckpt = torch.load(checkpoint, map_location='cpu', weights_only=False)
        sd = ckpt.get('g_ema', ckpt) if isinstance(ckpt, dict) else ckpt
        if m['baggan']:
            from .baggan import generator_from_baggan
            return generator_from_baggan(sd, img_resolution=m['size']).to(device)
        gen = Generator(m['size'], 512, 8)
        gen.load_state_dict(sd, strict=False)
        return gen.to(device)
    torch.manual_seed(seed)
    if m['baggan']:
        from .baggan import baggan_channels
        return Generator(m['size'], 512, 8, channels=baggan_channels()).to(device)
    return Generator(m['size'], 512, 8).to(device)
