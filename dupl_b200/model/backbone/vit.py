"""ViT-B/16 parameter container with the reference's state-dict schema.

Reference: model/backbone/vit.py:87-184 (Mlp, Attention, Block, PatchEmbed), :223-334
(VisionTransformer).  The modules below only OWN parameters (same names, shapes and
initialisation as the reference, so released checkpoints load with strict=True); the forward
pass is not executed by torch.nn — dupl_b200.encoder sequences the CUDA kernels over them.
"""
import torch
import torch.nn as nn

PATCH = 16


def _ln(dim):
    return nn.LayerNorm(dim, eps=1e-6)  # deit.py:100


class _Attn(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.scale = (dim // heads) ** -0.5
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, heads, ratio):
        super().__init__()
        self.norm1 = _ln(dim)
        self.attn = _Attn(dim, heads)
        self.norm2 = _ln(dim)
        self.mlp = _Mlp(dim, int(dim * ratio))


class _PatchEmbed(nn.Module):
    def __init__(self, img_size, dim):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (PATCH, PATCH)
        self.num_patches = (img_size // PATCH) ** 2
        self.proj = nn.Conv2d(3, dim, kernel_size=PATCH, stride=PATCH)


class VisionTransformer(nn.Module):
    def __init__(self, img_size=224, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0, num_classes=1000,
                 aux_layer=-3):
        super().__init__()
        if (embed_dim, depth, num_heads) != (768, 12, 12):
            raise NotImplementedError("libdupl.so is built for ViT-B/16 (768 wide, 12 blocks, 12 heads) only")
        self.embed_dim = self.num_features = embed_dim
        self.num_classes = num_classes
        self.patch_size = PATCH
        self.aux_layer = aux_layer
        self._size = img_size // PATCH
        self.patch_embed = _PatchEmbed(img_size, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim), requires_grad=False)
        self.blocks = nn.ModuleList(_Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth))
        self.norm = _ln(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes)  # unused by the path, kept for checkpoint compatibility
        nn.init.trunc_normal_(self.pos_embed, std=0.02, a=-2, b=2)
        nn.init.trunc_normal_(self.cls_token, std=0.02, a=-2, b=2)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02, a=-2, b=2)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def aux_block_index(self):
        """embeds[aux_layer] of vit.py:319-326 as a 0-based block index (11 == final-normed tokens)."""
        return self.aux_layer % len(self.blocks)

    def forward(self, *a, **k):
        raise RuntimeError("the encoder is driven by dupl_b200.encoder (CUDA kernels), not by nn.Module.forward")


DEIT_BASE_URL = "https://dl.fbaipublicfiles.com/deit/deit_base_patch16_224-b5f2ef4d.pth"


def deit_base_patch16_224(pretrained=False, **kwargs):
    """model/backbone/deit.py:97-109: with pretrained=True (the scripts' default, train_final_voc.py:54) the DeiT-B
    checkpoint comes through torch.hub exactly as in the reference — from the hub cache `./pretrained/` when the file is
    there (no network needed), else torch.hub tries to download it and raises what the reference would raise offline."""
    model = VisionTransformer(**kwargs)
    if pretrained:
        checkpoint = torch.hub.load_state_dict_from_url(url=DEIT_BASE_URL, model_dir="./pretrained", map_location="cpu",
                                                        check_hash=True)["model"]
        model.load_state_dict(checkpoint)
    return model


def vit_base_patch16_224(pretrained=False, **kwargs):
    if pretrained:
        raise RuntimeError("vit_base_patch16_224(pretrained=True) goes through timm's load_pretrained in the reference "
                           "(vit.py:1069-1080), which this image lacks: load the checkpoint with load_state_dict instead")
    return VisionTransformer(**kwargs)
