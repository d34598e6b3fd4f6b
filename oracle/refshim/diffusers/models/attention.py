def _chunked_feed_forward(ff, hidden_states, chunk_dim, chunk_size):
    import torch
    num_chunks = hidden_states.shape[chunk_dim] // chunk_size
    return torch.cat([ff(h) for h in hidden_states.chunk(num_chunks, dim=chunk_dim)], dim=chunk_dim)
