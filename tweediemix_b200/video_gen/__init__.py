"""Video stage (BASELINE configs[4]): the two things TweedieMix adds to diffusers' I2VGen-XL pipeline — the frame-0
residual-feature injection hooks (``utils_attn``) and the v-prediction Tweedie / DDIM step (``pipeline_step``) — on the tmx kernels.
The I2VGen-XL U-Net itself (3-D convolutions, temporal attention, image encoders) is the reference's unmodified diffusers dependency and
is not rebuilt here."""
