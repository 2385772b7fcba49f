"""CLAP2Diffusion Gradio application on B200 -- drop-in for the reference's ``app/gradio_app.py``.

Same surface: ``AudioToImageGenerator(checkpoint_dir).generate(audio_path, text_prompt, norm_value, num_steps, cfg_scale,
seed, model_type) -> (uint8 image [512,512,3], info string)`` (reference app/gradio_app.py:21-92), the same widgets
and launch environment variables (:95-196) -- but ``generate`` runs the real pipeline of ``scripts/inference.py``
(CLAP tower -> hierarchical tokens -> audio attention processors on the SD-1.5 UNet -> DDIM -> VAE) on libc2d instead
of returning ``np.random.randn(512, 512, 3)`` (:76-77).  ``gradio`` is imported only by ``build_demo`` / ``main`` (it is
not installed on the build box); the generator class works without it.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from pathlib import Path

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.append(ROOT)


def _inference_module():
    spec = importlib.util.spec_from_file_location("c2d_inference", os.path.join(ROOT, "scripts", "inference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class AudioToImageGenerator:
    """Main pipeline for audio-to-image generation (reference :21-92)."""

    MODEL_TYPES = ("Hierarchical", "SonicDiffusion", "Baseline")

    def __init__(self, checkpoint_dir="checkpoints"):
        self.checkpoint_dir = Path(checkpoint_dir)
        self.load_models()

    def load_models(self):
        """Checkpoints are resolved by scripts/inference.py (same file names as the reference reads, :38-47)."""
        self.inference = _inference_module().AudioToImageInference(checkpoint_dir=str(self.checkpoint_dir))
        self.device = self.inference.device

    def generate(self, audio_path, text_prompt, norm_value=60, num_steps=50, cfg_scale=7.5, seed=-1, model_type="Hierarchical"):
        """Returns (image uint8 [H,W,3], info).  model_type: "Hierarchical" conditions on the audio (when trained audio
        weights are loaded), "Baseline" is the text-only image; "SonicDiffusion" is a label of the reference's demo with no
        model behind it and maps to "Hierarchical"."""
        if seed is None or int(seed) == -1:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        seed = int(seed)
        if audio_path is None:
            raise ValueError("no audio given")
        self.inference.OPTIMAL_NORM = float(norm_value)
        image = self.inference.generate(audio_path, text_prompt or "", num_inference_steps=int(num_steps), guidance_scale=float(cfg_scale),
                                        seed=seed, use_hierarchical=(model_type != "Baseline"))
        info = f"""
Generation Complete!
Model: {model_type}
Audio: {Path(str(audio_path)).name if audio_path else 'None'}
Text: {text_prompt}
Norm: {norm_value}
Steps: {num_steps}
CFG: {cfg_scale}
Seed: {seed}
Audio conditioning: {'on' if (self.inference.audio_conditioning and model_type != 'Baseline') else 'off'}
"""
        return np.asarray(image), info


def build_demo(generator: AudioToImageGenerator):
    """The reference's Blocks layout (:98-175)."""
    import gradio as gr
    with gr.Blocks(title="CLAP2Diffusion - Audio to Image Generation") as demo:
        gr.Markdown("# CLAP2Diffusion: Audio-to-Image Generation\n### Hierarchical Audio Processing with Norm Optimization")
        with gr.Row():
            with gr.Column(scale=1):
                audio_input = gr.Audio(label="Upload Audio", type="filepath", sources="upload")
                model_dropdown = gr.Dropdown(choices=list(AudioToImageGenerator.MODEL_TYPES), value="Hierarchical", label="Model Type")
                text_input = gr.Textbox(label="Text Prompt", placeholder="Enter a description...", value="a beautiful landscape")
                with gr.Accordion("Advanced Settings", open=False):
                    norm_slider = gr.Slider(minimum=10, maximum=200, value=60, step=5, label="Audio Normalization (60 is optimal)")
                    steps_slider = gr.Slider(minimum=20, maximum=100, value=50, step=5, label="Inference Steps")
                    cfg_slider = gr.Slider(minimum=1, maximum=20, value=7.5, step=0.5, label="CFG Scale")
                    seed_input = gr.Number(label="Seed (-1 for random)", value=-1, precision=0)
                generate_btn = gr.Button("Generate Image", variant="primary")
            with gr.Column(scale=1):
                output_image = gr.Image(label="Generated Image")
                output_info = gr.Textbox(label="Generation Info", lines=9)
        generate_btn.click(fn=generator.generate,
                           inputs=[audio_input, text_input, norm_slider, steps_slider, cfg_slider, seed_input, model_dropdown],
                           outputs=[output_image, output_info])
    return demo


def main():
    generator = AudioToImageGenerator(os.getenv("C2D_CHECKPOINT_DIR", "checkpoints"))
    demo = build_demo(generator)
    auth_user, auth_pass = os.getenv("GRADIO_USERNAME", "admin"), os.getenv("GRADIO_PASSWORD", "clap2diffusion")
    demo.launch(share=False, server_name=os.getenv("GRADIO_SERVER_NAME", "127.0.0.1"),
                server_port=int(os.getenv("GRADIO_SERVER_PORT", 7860)), auth=(auth_user, auth_pass) if auth_pass else None)


if __name__ == "__main__":
    main()
