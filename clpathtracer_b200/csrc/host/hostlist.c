/* hostlist.c -- fat-pointer byte vectors (clpt_host.h "lists").
 *
 * Behavioural contract follows the reference's src/list.c:27-111: a header of
 * {capacity, length} (bytes) precedes the data pointer handed to the caller;
 * list_grow() returns old_length/bytes so vector_append() can index the new
 * slot; growth policy is capacity*2 + length + bytes (list.c:97-99).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "clpt_host.h"

size_t LIST_INDEX;

typedef struct list_hdr {
    size_t capacity;
    size_t length;
} list_hdr;

#define HDR(p) (((list_hdr *)(p)) - 1)
#define CHDR(p) (((const list_hdr *)(p)) - 1)

static list_hdr *
hdr_alloc(size_t payload) {
    list_hdr *h = malloc(sizeof(list_hdr) + payload);
    if (h == NULL) {
        perror("malloc");
        exit(EXIT_FAILURE);
    }
    return h;
}

static list_hdr *
hdr_resize(list_hdr *h, size_t payload) {
    list_hdr *n = realloc(h, sizeof(list_hdr) + payload);
    if (n == NULL) {
        free(h);
        perror("realloc");
        exit(EXIT_FAILURE);
    }
    n->capacity = payload;
    return n;
}

void *
new_list(size_t capacity_bytes) {
    list_hdr *h = hdr_alloc(capacity_bytes);
    h->capacity = capacity_bytes;
    h->length = 0;
    return h + 1;
}

void *
init_list(size_t count, size_t elem_size) {
    size_t bytes = count * elem_size;
    list_hdr *h = hdr_alloc(bytes);
    h->capacity = bytes;
    h->length = bytes;
    return h + 1;
}

void *
copy_list(const void *list) {
    size_t bytes = CHDR(list)->length;
    void *dup = init_list(1, bytes);
    memcpy(dup, list, bytes);
    return dup;
}

void
delete_list(void *list) {
    if (list != NULL) {
        free(HDR(list));
    }
}

size_t
list_size(const void *list) {
    return CHDR(list)->length;
}

size_t
list_grow(void **list_ptr, size_t bytes) {
    list_hdr *h = HDR(*list_ptr);
    size_t need = h->length + bytes;
    if (need > h->capacity) {
        h = hdr_resize(h, h->capacity * 2 + need);
    }
    h->length = need;
    *list_ptr = h + 1;
    return need / bytes - 1;
}

void
list_concat(void **list1_ptr, const void *list2) {
    list_hdr *a = HDR(*list1_ptr);
    size_t na = a->length, nb = CHDR(list2)->length;
    a = hdr_resize(a, na + nb);
    memcpy((char *)(a + 1) + na, list2, nb);
    a->length = na + nb;
    *list1_ptr = a + 1;
}
