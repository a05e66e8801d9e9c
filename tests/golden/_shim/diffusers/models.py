class AutoencoderKL:  # only used as a type annotation by the reference schedulers
    pass
