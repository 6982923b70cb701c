/* emu_switch.S -- fiber context switch of the test-only SIMT emulator (tests/emu/cuda_emu.h), x86-64 SysV.
 * void pna_emu_switch(void** save_sp, void* load_sp): pushes the callee-saved registers, stores rsp, loads the other
 * fiber's rsp, pops its callee-saved registers and returns into it. */
    .text
    .globl pna_emu_switch
    .type pna_emu_switch,@function
pna_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size pna_emu_switch, .-pna_emu_switch
    .section .note.GNU-stack,"",@progbits
