"""Stub for mmgp==3.4.9 (TEST INFRASTRUCTURE ONLY): the reference only needs
offload.shared_state (attention backend switch) on the bf16/fp32 arithmetic path."""


class _Offload:
    shared_state = {"_attention": "sdpa"}
    last_offload_obj = None

    @staticmethod
    def set_step_no_for_lora(*args, **kwargs):
        pass


offload = _Offload()
profile_type = None
