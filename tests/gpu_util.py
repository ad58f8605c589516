import torch


def relerr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-12))


def cosine(a, b):
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp(min=1e-30))


DEV = "cuda:0"
