"""Evaluation script for CLAP2Diffusion on B200 -- drop-in for the reference's ``scripts/evaluate.py``.

Same ``Evaluator`` surface (evaluate_single / evaluate_dataset / print_results, reference :19-146), the same CLI
(``--data_dir --checkpoint_dir --output_dir``, :148-179), the same dataset layout (``metadata.json`` = list of
{"audio", "text", "id"}, clips under ``audio/``) and the same outputs (``<id>_generated.png`` per item and
``evaluation_results.json`` with ``individual_results`` / ``average_metrics``).  The images come from the real pipeline
(scripts/inference.py on libc2d).

Metrics: the reference's ``compute_clip_score`` / ``compute_audio_alignment`` are random-number placeholders (:32-40).
A CLIP image / text encoder is not part of this repository and cannot be downloaded on the box, so both return ``None``
("not measured") unless the caller provides scorers (``Evaluator(..., clip_scorer=f, audio_scorer=g)`` with
``f(image: PIL.Image, text: str) -> float`` and ``g(image, audio_embedding [1,512]) -> float``); what IS always measured is
the wall time per image and basic image statistics, so a run is never silently decorated with made-up scores.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.append(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _inference_module():
    spec = importlib.util.spec_from_file_location("c2d_inference", os.path.join(os.path.dirname(os.path.abspath(__file__)), "inference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class Evaluator:
    def __init__(self, checkpoint_dir="../checkpoints", clip_scorer=None, audio_scorer=None, num_inference_steps=50):
        self.checkpoint_dir = Path(checkpoint_dir)
        self.pipeline = _inference_module().AudioToImageInference(str(checkpoint_dir))
        self.clip_scorer, self.audio_scorer = clip_scorer, audio_scorer
        self.num_inference_steps = num_inference_steps
        self.metrics = {"clip_score": [], "fid_score": [], "inception_score": [], "audio_alignment": [], "seconds_per_image": []}

    def compute_clip_score(self, image, text):
        return None if self.clip_scorer is None else float(self.clip_scorer(image, text))

    def compute_audio_alignment(self, image, audio_embedding):
        return None if self.audio_scorer is None else float(self.audio_scorer(image, audio_embedding))

    def evaluate_single(self, audio_path, text_prompt, reference_image=None):
        t0 = time.time()
        generated_image = self.pipeline.generate(audio_path=audio_path, text_prompt=text_prompt, seed=42,      # fixed seed (:49)
                                                 num_inference_steps=self.num_inference_steps)
        seconds = time.time() - t0
        audio_embedding = self.pipeline.extract_clap_embedding(self.pipeline.load_audio(audio_path))
        a = np.asarray(generated_image).astype(np.float32)
        return {"clip_score": self.compute_clip_score(generated_image, text_prompt),
                "audio_alignment": self.compute_audio_alignment(generated_image, audio_embedding),
                "generated_image": generated_image, "seconds": seconds,
                "image_mean": float(a.mean()), "image_std": float(a.std())}

    def evaluate_dataset(self, data_dir, output_dir="evaluation_results"):
        data_dir, output_dir = Path(data_dir), Path(output_dir)
        output_dir.mkdir(parents=True, exist_ok=True)
        metadata_path = data_dir / "metadata.json"
        if metadata_path.exists():
            with open(metadata_path, "r") as f:
                metadata = json.load(f)
        else:
            metadata = [{"audio": "sample1.wav", "text": "thunder and rain", "id": "001"},
                        {"audio": "sample2.wav", "text": "birds chirping", "id": "002"}]
        print(f"Evaluating {len(metadata)} samples...")
        all_results = []
        for item in metadata:
            audio = str(item["audio"])
            audio_path = audio if audio.startswith("synthetic:") else data_dir / "audio" / audio
            if not audio.startswith("synthetic:") and not Path(audio_path).exists():
                print(f"Warning: {audio_path} not found, skipping...")
                continue
            results = self.evaluate_single(audio_path=audio_path, text_prompt=item["text"])
            for k in ("clip_score", "audio_alignment"):
                if results[k] is not None:
                    self.metrics[k].append(results[k])
            self.metrics["seconds_per_image"].append(results["seconds"])
            results["generated_image"].save(output_dir / f"{item['id']}_generated.png")
            all_results.append({"id": item["id"], "audio": item["audio"], "text": item["text"], "clip_score": results["clip_score"],
                                "audio_alignment": results["audio_alignment"], "seconds": results["seconds"],
                                "image_mean": results["image_mean"], "image_std": results["image_std"]})
        avg_metrics = {}
        for name, values in self.metrics.items():
            if values:
                avg_metrics[name] = float(np.mean(values))
                avg_metrics[f"{name}_std"] = float(np.std(values))
        with open(output_dir / "evaluation_results.json", "w") as f:
            json.dump({"individual_results": all_results, "average_metrics": avg_metrics,
                       "not_measured": [k for k in ("clip_score", "audio_alignment", "fid_score", "inception_score") if not self.metrics[k]],
                       "audio_conditioning": bool(self.pipeline.audio_conditioning)}, f, indent=2)
        return avg_metrics

    def print_results(self, metrics):
        print("\n" + "=" * 60)
        print("Evaluation Results")
        print("=" * 60)
        for name, value in metrics.items():
            if not name.endswith("_std"):
                std_key = f"{name}_std"
                print(f"{name:20}: {value:.4f} ± {metrics[std_key]:.4f}" if std_key in metrics else f"{name:20}: {value:.4f}")


def main(argv=None):
    parser = argparse.ArgumentParser(description="Evaluate CLAP2Diffusion")
    parser.add_argument("--data_dir", type=str, default="../data/audiocaps/test", help="Path to test data directory")
    parser.add_argument("--checkpoint_dir", type=str, default="../checkpoints", help="Path to checkpoint directory")
    parser.add_argument("--output_dir", type=str, default="evaluation_results", help="Output directory for results")
    args = parser.parse_args(argv)
    print("\n" + "=" * 60)
    print("CLAP2Diffusion Evaluation")
    print("=" * 60)
    print(f"\nData directory: {args.data_dir}\nCheckpoint directory: {args.checkpoint_dir}\nOutput directory: {args.output_dir}\n")
    evaluator = Evaluator(checkpoint_dir=args.checkpoint_dir)
    metrics = evaluator.evaluate_dataset(data_dir=args.data_dir, output_dir=args.output_dir)
    evaluator.print_results(metrics)
    print(f"\n✓ Evaluation complete! Results saved to {args.output_dir}")


if __name__ == "__main__":
    main()
