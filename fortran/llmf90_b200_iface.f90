! llmf90_b200_iface.f90 -- ISO_C_BINDING interface to libllmf90_b200.so (include/llmf90_b200.h).
!
! This module is what a maintainer of llm.f90 adds next to weight_module.f90 to route
! `transformer(token,pos,s,w)` (llama2.f90:480-640) through the B200 library.  It follows the
! bind(C) convention of the author's own helper interface (load.f90:123-152: scalar arguments by
! `value`, c_float / c_int kinds).  It was desk-checked only: no Fortran compiler exists in the
! image this repository is built in (see INTEGRATION.md for the three call-site edits).
module llmf90_b200
        use iso_c_binding
        implicit none

        integer(c_int32_t), parameter :: LLMF90_WTYPE_F32 = 0, LLMF90_WTYPE_F16 = 1, LLMF90_WTYPE_Q4_0 = 2
        ! flags: granular kernels, per-phase timers, batched prompt pass, Q6_K output.weight
        integer(c_int32_t), parameter :: LLMF90_FLAG_GRANULAR = 1, LLMF90_FLAG_PROFILE = 2, LLMF90_FLAG_PREFILL = 4, &
                & LLMF90_FLAG_CLS_Q6K = 8

        ! mirror of `llmf90_b200_config`; the first seven fields are type Config (weight_module.f90:28-31)
        type, bind(C) :: b200_config
                integer(c_int32_t) :: emb_dim, hidden_dim, n_layers, n_heads, n_kv_heads, vocab_size, seq_len
                integer(c_int32_t) :: wtype = 0
                integer(c_int32_t) :: device = 0
                integer(c_int32_t) :: tp_rank = 0
                integer(c_int32_t) :: tp_size = 1
                integer(c_int32_t) :: flags = 0
        end type b200_config

        interface
                ! replaces the RunState allocation (llama2.f90:311-319); copies the weights to the GPU
                function b200_init(cfg, token_embedding_table, rms_att_weight, wqkv, wo, rms_ffn_weight, &
                                & w13, w2, rms_final_weight, wcls) bind(C, name="llmf90_b200_init") result(rc)
                        import :: c_int, c_float, b200_config
                        type(b200_config), intent(in) :: cfg
                        real(c_float), intent(in) :: token_embedding_table(*), rms_att_weight(*), wqkv(*), wo(*)
                        real(c_float), intent(in) :: rms_ffn_weight(*), w13(*), w2(*), rms_final_weight(*), wcls(*)
                        integer(c_int) :: rc
                end function b200_init

                ! logits = transformer(token,pos,s,w)  (llama2.f90:380); token and pos are the 1-based
                ! values the Fortran loop already holds
                function b200_transformer(token, pos, logits) bind(C, name="llmf90_b200_transformer") result(rc)
                        import :: c_int, c_int32_t, c_float
                        integer(c_int32_t), value :: token, pos
                        real(c_float), intent(out) :: logits(*)
                        integer(c_int) :: rc
                end function b200_transformer

                ! s%times(1:5) in milliseconds (llama2.f90:407-410)
                function b200_times(t) bind(C, name="llmf90_b200_times") result(rc)
                        import :: c_int, c_float
                        real(c_float), intent(out) :: t(5)
                        integer(c_int) :: rc
                end function b200_times

                function b200_reset() bind(C, name="llmf90_b200_reset") result(rc)
                        import :: c_int
                        integer(c_int) :: rc
                end function b200_reset

                function b200_free() bind(C, name="llmf90_b200_free") result(rc)
                        import :: c_int
                        integer(c_int) :: rc
                end function b200_free

                function b200_last_error() bind(C, name="llmf90_b200_last_error") result(msg)
                        import :: c_ptr
                        type(c_ptr) :: msg
                end function b200_last_error

                ! the inner subroutines as separately callable operators
                function b200_matvec(w, wtype, rows, cols, x, y) bind(C, name="llmf90_b200_matvec") result(rc)
                        import :: c_int, c_int32_t, c_float
                        real(c_float), intent(in) :: w(*), x(*)
                        integer(c_int32_t), value :: wtype, rows, cols
                        real(c_float), intent(out) :: y(*)
                        integer(c_int) :: rc
                end function b200_matvec

                function b200_rmsnorm(x, w, n, xr) bind(C, name="llmf90_b200_rmsnorm") result(rc)
                        import :: c_int, c_int32_t, c_float
                        real(c_float), intent(in) :: x(*), w(*)
                        integer(c_int32_t), value :: n
                        real(c_float), intent(out) :: xr(*)
                        integer(c_int) :: rc
                end function b200_rmsnorm

                function b200_softmax(x, n, s, p) bind(C, name="llmf90_b200_softmax") result(rc)
                        import :: c_int, c_int32_t, c_float
                        real(c_float), intent(in) :: x(*)
                        integer(c_int32_t), value :: n, s
                        real(c_float), intent(out) :: p(*)
                        integer(c_int) :: rc
                end function b200_softmax

                function b200_rope(q, k, emb, kv, head_size, pos) bind(C, name="llmf90_b200_rope") result(rc)
                        import :: c_int, c_int32_t, c_float
                        real(c_float), intent(inout) :: q(*), k(*)
                        integer(c_int32_t), value :: emb, kv, head_size, pos
                        integer(c_int) :: rc
                end function b200_rope

                ! the forced prompt positions (llama2.f90:379-385) as one batched pass: fills the KV cache for the
                ! input tokens at positions pos0 .. pos0 + n_tokens - 1 (needs LLMF90_FLAG_PREFILL)
                function b200_prefill(tokens, n_tokens, pos0) bind(C, name="llmf90_b200_prefill") result(rc)
                        import :: c_int, c_int32_t
                        integer(c_int32_t), intent(in) :: tokens(*)
                        integer(c_int32_t), value :: n_tokens, pos0
                        integer(c_int) :: rc
                end function b200_prefill

                ! transformer() + maxloc (temperature == 0, llama2.f90:388) or softmax(logits / temperature) + the
                ! CDF walk of `sample` against r (llama2.f90:390-391, :428-447) on the device; r from random_number
                function b200_transformer_sample(token, pos, temperature, r, next_token) &
                                & bind(C, name="llmf90_b200_transformer_sample") result(rc)
                        import :: c_int, c_int32_t, c_float
                        integer(c_int32_t), value :: token, pos
                        real(c_float), value :: temperature, r
                        integer(c_int32_t), intent(out) :: next_token
                        integer(c_int) :: rc
                end function b200_transformer_sample
        end interface

contains

        ! print the library's message and stop, the reference's error convention (read_ggml.f90:122-125)
        subroutine b200_check(rc)
                integer(c_int), intent(in) :: rc
                character(kind=c_char), pointer :: s(:)
                integer :: n
                if (rc == 0) return
                call c_f_pointer(b200_last_error(), s, [1024])
                n = 0
                do while (n < 1024)
                        if (s(n + 1) == c_null_char) exit
                        n = n + 1
                end do
                print *, "llmf90_b200: ", s(1:n)
                stop 1
        end subroutine b200_check

end module llmf90_b200
